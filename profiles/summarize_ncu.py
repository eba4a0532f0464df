"""Turn an `ncu --set full` report and an ncu launch list into the summaries committed next to this file.

    ncu -i gpurun_out/r1c_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py full /tmp/raw.csv profiles/r1c_ncu_full_summary.json profiles/r1c_traffic.json
    python profiles/summarize_ncu.py launches gpurun_out/r1c_launches.csv profiles/r1c_launches_summary.txt

The captures themselves are made on a GPU box (see profiles/README.md for the command lines); the .ncu-rep files are too
large to commit and stay in gpurun_out/.
"""
import collections
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]
NAMES = {"photo_kernel<32, 32, 0": "photo_jac", "photo_kernel<32, 32, 1": "photo_err", "geo_kernel<32, 1>": "geo_jac",
         "geo_tc_kernel<32>": "geo_jac", "geo_kernel<32, 0>": "geo_err"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def full(raw_csv, out_json, traffic_json):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units = rows[0], rows[1]
    out = {}
    traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture summarised in "
                           + out_json.split("/")[-1] + " (180 pairs per launch, N=1)"}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {k: [r[hdr.index(k)], units[hdr.index(k)]] for k in KEYS if k in hdr}
        st = {}
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued"):
                try:
                    if float(r[i]) > 0:
                        st[h.split("stalled_")[1]] = float(r[i])
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda t: -t[1])}
        out[name] = d
        for kk, vv in NAMES.items():
            if kk in name:
                rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
                traffic[vv] = {"dram_bytes_per_launch": float(rd[0]) * UNIT[rd[1]] + float(wr[0]) * UNIT[wr[1]], "pairs_per_launch": 180}
    json.dump(out, open(out_json, "w"), indent=1)
    json.dump(traffic, open(traffic_json, "w"), indent=1)


def launches(launch_csv, out_txt, header=""):
    rows = list(csv.reader(l for l in open(launch_csv) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        a = agg.setdefault(r[ik].split("(")[0][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e6  # ns -> ms
    tot = sum(a[1] for a in agg.values())
    lines = [header] if header else []
    for k, (n, v) in sorted(agg.items(), key=lambda t: -t[1][1])[:22]:
        lines.append(f"{v:9.3f} ms {n:4d} launches {100 * v / tot:5.1f}%  {k}")
    lines.append(f"{tot:9.3f} ms total")
    open(out_txt, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(*sys.argv[2:5])
    else:
        launches(*sys.argv[2:4])

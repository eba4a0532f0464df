#!/usr/bin/env python
"""bench.py -- LM iterations/s and residuals/s of the 32-keyframe local BA (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]

Workload (BASELINE.json configs[3], SURVEY.md section 8d "Config 4"): 32 keyframes, 320x256 maps,
F = 32 feature channels, C = 32 code entries, L = 4 pyramid levels, dense sampling (N = 81920);
factor graph = temporal links with 3 back-connections in both directions (180 ordered pairs), each with a
photometric, a geometric and a reprojection (M = 512) factor, plus code / scale priors; KF0 anchors the gauge.
One step = one LM iteration: linearise every factor -> [all-reduce] -> assemble -> Schur solve -> evaluate the
candidate -> [all-reduce] -> accept / reject.  The problem is fixed as N grows ("strong" scaling): ordered
pairs are sharded round-robin over the ranks, one NCCL all-reduce of the packed factor buffer per iteration.

value   : LM iterations/s with everything resident in HBM (CUDA events on the context stream, max over ranks)
e2e     : the same iteration through the public API with the state coming from / going to pinned HOST memory
          every step (set_state H2D, get_state + cost D2H inside the timed region).  The keyframe maps stay on
          the device, as they do in the reference (Frame tensors are CUDA tensors, core/mapping/frame.h).
roofline: the photometric linearisation kernel (the dominant launch), algorithmic bytes of SURVEY.md 8(d) per
          launch / its CUDA-event duration, against MEASURED_PEAKS.json's hbm_gbs.
--impl reference: the reference has NO CPU implementation of this path, so the reference arm
          times the CPU restatement of its kernels (oracle/, all host threads) on a bounded sample (one ordered
          pair: photometric + geometric linearisation and error evaluation) and extrapolates to the 180 pairs.
--impl shim: the same pair-by-pair call sequence as reference-gpu, but through integration/df_sage_shim.cpp (the
reference's df:: symbols implemented by libsage_ba.so) -- what an unmodified caller of the reference gets.
--impl reference-gpu: the reference's OWN CUDA kernels (oracle/_ref, compiled unmodified from /root/reference),
          called pair by pair as core/gtsam/*_factor.cpp does, on the same B200 (bounded sample of pairs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(num_kf=32, W=320, H=256, L=4, F=32, C=32, back_connections=3, matches=512)
PHOTO_W = [10.0, 9.0, 8.0, 7.0]
EPS = 1e-4
METRIC = "LM iterations/s, 32-KF local BA (320x256x32 feat, 32-dim code, photometric+geometric+reprojection)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is polled in-process every
    ~2 ms (a timed region of a few tens of ms would be over before an `nvidia-smi -lms` child has started); only samples
    taken between mark_begin() and mark_end() are reported.  Falls back to one nvidia-smi query if NVML is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, device=0):
        self.device, self.rows, self.run, self.t0, self.t1, self.h, self.nv = device, [], False, None, None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.run = True
        self.th = threading.Thread(target=self._poll, daemon=True)
        self.th.start()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                      str(self.device)], capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                        "note": "NVML unavailable: one nvidia-smi sample right after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.run = False
        self.th.join(timeout=1)
        rows = [r for r in self.rows if self.t0 is None or (self.t0 <= r[0] <= (self.t1 or r[0]))] or self.rows[-1:]
        sm = [r[1] for r in rows]
        mask = 0
        for r in rows:
            mask |= int(r[2])
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": [n for n, bit in self.REASONS if mask & bit], "samples": len(sm)}


def build_scene(wl):
    import sage_slam_b200 as sage

    kfs = sage.synthetic.make_scene(num_kf=wl["num_kf"], W=wl["W"], H=wl["H"], L=wl["L"], F=wl["F"], C=wl["C"],
                                    back_connections=wl["back_connections"], seed=1234)
    pairs = sage.synthetic.ordered_pairs(kfs)
    return kfs, pairs


def algorithmic_bytes(wl):
    """bytes one photometric / geometric linearisation of ONE ordered pair must read + write."""
    F, C, N = wl["F"], wl["C"], wl["W"] * wl["H"]
    P0 = wl["W"] * wl["H"]
    SP, w, h = 0, wl["W"], wl["H"]
    for _ in range(wl["L"]):
        SP += w * h
        w, h = w // 2, h // 2
    Dp, Dg = 13 + C, 14 + 2 * C
    photo = 4 * (4 * F * SP + N * (C + 5) + P0) + 4 * (Dp * Dp + Dp + 2)
    photo_err = 4 * (2 * F * SP + N * (C + 5) + P0) + 8
    geo = 4 * (P0 * (C + 4) + N * (C + 5)) + 4 * (Dg * Dg + Dg + 2)
    return photo, photo_err, geo


def add_factors(ba, kfs, pairs, wl, sage):
    geo_loss = float(0.03 * np.mean(kfs[0].dpt_map_bias.astype(np.float64) ** 2))
    for (i, j) in pairs:
        ba.add_photometric(i, j, PHOTO_W[:wl["L"]])
    for (i, j) in pairs:
        ba.add_geometric(i, j, geo_loss, 0.1)
    if wl.get("matches", 0):
        for (i, j) in pairs:
            loc, homo, uv = sage.synthetic.make_matches(kfs[i], kfs[j], M=wl["matches"])
            ba.add_reprojection(i, j, loc, homo, uv, 0.03 * wl["W"] ** 2, 0.1)
    for k in range(len(kfs)):
        ba.add_code_prior(k, 1e-3)
        ba.add_scale_prior(k, 1.0, 1e-2)
    ba.fix(0, pose=True, scale=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import sage_slam_b200 as sage

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout and would precede the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    wl = dict(WORKLOAD)
    if args.small:
        wl.update(num_kf=4, W=128, H=96, F=16, C=8, matches=64)
    kfs, pairs = build_scene(wl)
    stream = torch.cuda.Stream(device=local)
    with torch.cuda.stream(stream):
        ctx = sage.Context(local, stream=stream.cuda_stream)
        dkfs = [sage.DeviceKeyframe(ctx, k) for k in kfs]
        ba = sage.LocalBA(ctx, dkfs, rank=rank, world=world)
        add_factors(ba, kfs, pairs, wl, sage)
        poses0 = [k.pose_wk for k in kfs]
        codes0 = np.stack([k.code for k in kfs])
        scales0 = np.array([k.dpt_scale for k in kfs], np.float32)
        ba.set_state(poses0, codes0, scales0, EPS)

        state = {"damp": 1e-4, "cost": None}

        def lm_iteration():
            # one LM iteration through the library's own driver: linearise (+ all-reduce of the packed factor buffer when
            # world > 1) -> assemble -> solve -> evaluate the candidate -> accept / reject, one host synchronisation
            cost, cand, _, state["damp"] = ba.lm_step(state["damp"], min_damp=1e-6, max_damp=1e2)
            state["cost"] = min(cand, cost)
            return cost, cand

        def sync_all():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        costs = []
        for _ in range(args.warmup):
            costs.append(lm_iteration())
        # ---- device-resident timing -------------------------------------------------------------------
        ba.set_state(poses0, codes0, scales0, EPS)
        state["damp"] = 1e-4
        ba.profile(True)
        ba.profile_read(reset=True)
        l0 = ctx.launch_count
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        sync_all()
        sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            costs.append(lm_iteration())
        e1.record(stream)
        sync_all()
        sampler.mark_end()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count - l0
        prof = ba.profile_read(reset=True)
        ba.profile(False)
        # ---- end-to-end: state from / to pinned host memory every step -----------------------------------
        K, C = len(kfs), wl["C"]
        P = np.stack([np.concatenate([R.reshape(-1), t.reshape(-1)]) for R, t in poses0]).astype(np.float32)
        pin_in = [torch.from_numpy(x.copy()).pin_memory() for x in (P, codes0.astype(np.float32), scales0)]
        h2d = sum(x.numel() * 4 for x in pin_in)
        d2h = h2d + 8
        import ctypes

        def e2e_step():
            ctx.check(ctx.lib.sage_ba_problem_set_state(ba.h, ctypes.c_void_p(pin_in[0].data_ptr()), ctypes.c_void_p(pin_in[1].data_ptr()),
                                                        ctypes.c_void_p(pin_in[2].data_ptr()), EPS))
            lm_iteration()
            new_poses, new_codes, new_scales = ba.get_state()
            pin_in[0].copy_(torch.from_numpy(np.stack([np.concatenate([R.reshape(-1), t]) for R, t in new_poses])))
            pin_in[1].copy_(torch.from_numpy(new_codes))
            pin_in[2].copy_(torch.from_numpy(new_scales))

        state["damp"] = 1e-4
        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        sync_all()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            tt = torch.tensor([ms, e2e_ms], device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, e2e_ms = float(tt[0]), float(tt[1])
            ll = torch.tensor([launches], device=f"cuda:{local}")
            dist.all_reduce(ll)
            launches = int(ll[0])

        tracker = None
        if rank == 0 and world == 1 and not args.no_tracker:
            tracker = tracker_bench(ctx, sage, kfs, dkfs, wl)

    if rank == 0:
        hbm, src = load_peaks()
        b_photo, b_photo_err, b_geo = algorithmic_bytes(wl)
        shard = ba.shard_counts()
        pj_ms, pj_n = prof["photo_jac"]
        per_launch_ms = pj_ms / max(pj_n, 1)
        achieved = shard["photo"] * b_photo / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        traffic = None
        try:  # DRAM bytes of the same launch from the committed ncu --set full capture (scaled to this rank's pair count)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1c_traffic.json")))["photo_jac"]
            traffic = tj["dram_bytes_per_launch"] / tj["pairs_per_launch"] * shard["photo"]
        except Exception:
            pass
        residuals = ba.num_residuals
        ms_per_step = ms / args.steps
        line = {
            "metric": METRIC, "value": 1e3 / ms_per_step, "unit": "LM iters/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "32-KF local BA, 180 ordered pairs x (photometric + geometric + reprojection M=512), "
                                   "320x256, F=32, C=32, L=4, dense N=81920" if not args.small else "SMALL debug workload",
                       "keyframes": wl["num_kf"], "ordered_pairs": len(pairs), "factors_per_rank": shard,
                       "parallelism": f"pair-sharded x{world}, 1 all-reduce/iter",
                       "l2": "inputs (1.8 GB of keyframe maps per rank) exceed the 126 MB L2; no flush needed"},
            "mresiduals_per_s": residuals / (ms_per_step * 1e-3) / 1e6, "residuals_per_iter": residuals,
            "e2e": {"value": 1e3 / (e2e_ms / args.steps), "unit": "LM iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "photo_kernel<32,32,MAP_JAC> (photometric linearisation, all owned pairs per launch)",
                         "achieved": achieved, "peak": hbm, "peak_source": src, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "launch_ms": per_launch_ms, "algorithmic_bytes_per_launch": shard["photo"] * b_photo,
                         "note": "bound is HBM by algorithmic bytes; the measured limiter is the L1 data pipe (61 % busy, DRAM 8 %: "
                                 "profiles/r1c_ncu_full_summary.json), see profiles/README.md"},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "clocks": clocks,
            "lm_trace": [(float(a), float(b)) for a, b in costs[-args.steps:]][:6],
        }
        if tracker is not None:
            line["config2_tracker"] = tracker
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N=1 only: torchrun pins OMP_NUM_THREADS=1
            line["cpu_baseline"] = cpu_baseline(wl, kfs, pairs, len(pairs))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def tracker_bench(ctx, sage, kfs, dkfs, wl, reps=5):
    """BASELINE configs[1]: CameraTracker::TrackNewFrame, 1 reference KF vs 1 live frame at 320x256x32, photometric +
    reprojection (M = 256), 10 LM iterations -- latency of the whole call (pre-sampling + LM loop, host solve included)."""
    import torch

    from sage_slam_b200 import ops

    k0, k1 = kfs[1], kfs[0]  # reference keyframe, frame to track
    R0, t0 = k0.pose_wk
    R1, t1 = k1.pose_wk
    R10 = (R1.T @ R0).astype(np.float32)
    t10 = (R1.T @ (t0 - t1)).astype(np.float32)
    loc, homo, uv = sage.synthetic.make_matches(k0, k1, M=256)
    dpts = k0.dpt_map.reshape(-1)[loc].astype(np.float32)
    times, rep = [], None
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t_0 = time.perf_counter()
        _, _, rep = ops.track_new_frame(ctx, dkfs[1], dkfs[0], k0.code, k0.dpt_scale, R10, t10, PHOTO_W[:wl["L"]], dpt_eps=EPS,
                                        max_num_iters=10, matches=(dpts, homo, uv), reproj_loss_param=0.03 * wl["W"] ** 2,
                                        reproj_weight=0.1)
        torch.cuda.synchronize()
        if r:
            times.append((time.perf_counter() - t_0) * 1e3)
    return {"workload": "TrackNewFrame: 1 KF vs 1 frame, 320x256, F=32, L=4, dense N=81920, photometric + reprojection M=256, max 10 LM iters",
            "ms_per_track": float(np.median(times)), "lm_iterations": rep["iterations"], "jacobian_evals": rep["jacobian_evals"],
            "error_evals": rep["error_evals"], "final_error": rep["final_error"]}


def cpu_sample(kfs, pair, wl):
    """One ordered pair through the CPU oracle: photometric + geometric linearisation and error evaluation."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import oracle as O

    a = helpers.case_args(kfs, pair[0], pair[1])
    t0 = time.perf_counter()
    O.photometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"], a["mask1"],
                            a["loc1d"], a["homo"], a["feat0"], a["feat1"], a["grad1"], a["level_offsets"], a["scale0"], a["cams"],
                            a["eps"], a["weights"])
    O.geometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"], a["dpt1"],
                          a["dgrad1"], a["basis1"], a["mask1"], a["loc1d"], a["homo"], a["scale0"], a["scale1"], a["cam"], a["eps"],
                          a["geo_loss"], a["geo_weight"])
    O.photometric_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], a["mask1"], a["loc1d"], a["homo"], a["feat0"],
                        a["feat1"], a["level_offsets"], a["scale0"], a["cams"], a["eps"], a["weights"])
    O.geometric_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], a["dpt1"], a["mask1"], a["loc1d"], a["homo"],
                      a["scale0"], a["cam"], a["eps"], a["geo_loss"], a["geo_weight"])
    return time.perf_counter() - t0, O.num_threads()


def cpu_baseline(wl, kfs, pairs, npairs, reps=1):
    ts = []
    for r in range(reps):
        t, cores = cpu_sample(kfs, pairs[r % len(pairs)], wl)
        ts.append(t)
    t_pair = float(np.mean(ts))
    return {"value": 1.0 / (t_pair * npairs), "unit": "LM iters/s", "cores": cores, "kind": "port",
            "sample": f"{reps} of {npairs} ordered pairs (photometric+geometric linearisation + error evaluation), "
                      f"{t_pair:.2f} s/pair, extrapolated to the full iteration; solve not included"}


def run_reference_cpu(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = dict(WORKLOAD)
    if args.small:
        wl.update(num_kf=4, W=128, H=96, F=16, C=8, matches=64)
    # only the keyframes the sampled pairs touch are needed
    sub = dict(wl)
    sub["num_kf"] = 4
    kfs, pairs = build_scene(sub)
    npairs = 180 if not args.small else len(pairs)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(kfs, pairs[0], wl)
    ts = []
    for s in range(args.steps):
        t, cores = cpu_sample(kfs, pairs[s % len(pairs)], wl)
        ts.append(t)
    t_pair = float(np.mean(ts))
    value = 1.0 / (t_pair * npairs)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "LM iters/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_pair * npairs * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "32-KF local BA, 180 ordered pairs (CPU restatement of the reference kernels; the reference "
                                   "itself has no CPU path, SURVEY.md fact 1)"},
            "cpu_baseline": {"value": value, "unit": "LM iters/s", "cores": cores, "kind": "port",
                             "sample": f"each step = 1 of {npairs} ordered pairs (photometric+geometric linearisation + error "
                                       f"evaluation), {t_pair:.2f} s/pair, extrapolated x{npairs}"},
            "e2e": {"value": value, "unit": "LM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_reference_gpu(args):
    """The reference's own CUDA kernels (oracle/_ref) pair by pair, as core/gtsam/*_factor.cpp drives them."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import build_ref
    import helpers

    wl = dict(WORKLOAD)
    if args.small:
        wl.update(num_kf=4, W=128, H=96, F=16, C=8, matches=64)
    sub = dict(wl)
    sub["num_kf"] = 4
    if args.ref_samples:
        sub["num_samples"] = args.ref_samples
    import sage_slam_b200 as sage

    kfs = sage.synthetic.make_scene(num_kf=4, W=wl["W"], H=wl["H"], L=wl["L"], F=wl["F"], C=wl["C"], back_connections=3, seed=1234,
                                    num_samples=args.ref_samples or None)
    pairs = sage.synthetic.ordered_pairs(kfs)
    npairs = 180 if not args.small else len(pairs)
    mod = build_ref.load_shim(wl["C"], wl["F"]) if args.impl == "shim" else build_ref.load(wl["C"], wl["F"])
    dev = torch.device("cuda:0")

    def T(a, dtype=torch.float32):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dtype)

    def prep(pair):
        a = helpers.case_args(kfs, pair[0], pair[1])
        R = {k: T(a[k]) for k in ("R10", "t10", "R0", "t0", "R1", "t1", "bias0", "code0", "code1", "mask1", "homo", "feat0", "feat1",
                                  "grad1")}
        R["jac0"] = T(np.ascontiguousarray(a["jac0"].T)).t()
        b = kfs[pair[1]]
        R["bias1"] = T(b.dpt_map_bias)
        R["jac1"] = T(np.ascontiguousarray(b.dpt_jac_code.T)).t()
        R["loc64"] = T(a["loc1d"], torch.int64)
        R["lo"] = T(a["level_offsets"], torch.int32)
        return a, R

    def one_pair(a, R):
        cam = [float(x) for x in a["cam"]]
        w = torch.tensor(a["weights"])
        H, W, C = a["H"], a["W"], a["C"]
        mod.photometric_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], R["jac0"], R["code0"], R["mask1"],
                                  R["loc64"], R["homo"], R["feat0"], R["feat1"], R["grad1"], R["lo"], a["scale0"], cam, a["L"],
                                  a["eps"], w)
        # GeometricFactor::ComputeJacobianAndError's per-call preparation (geometric_factor.cpp:317-320,340-342)
        un = R["bias1"].reshape(H, W) + torch.matmul(R["jac1"], R["code1"]).reshape(H, W)
        p = torch.nn.functional.pad(un.reshape(1, 1, H, W), (1, 1, 1, 1), mode="replicate")
        gx = 0.5 * (p[:, :, 1:H + 1, 2:W + 2] - p[:, :, 1:H + 1, 0:W])
        gy = 0.5 * (p[:, :, 2:H + 2, 1:W + 1] - p[:, :, 0:H, 1:W + 1])
        grad = torch.cat([gx, gy], 1)
        mod.geometric_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], R["jac0"], R["code0"],
                                a["scale1"] * un, a["scale1"] * grad.reshape(2, H, W), R["jac1"].reshape(H, W, C), R["mask1"],
                                R["loc64"].to(torch.int32), R["homo"], a["scale0"], a["scale1"], cam, a["eps"], a["geo_loss"],
                                a["geo_weight"])
        mod.photometric_error(R["R10"], R["t10"], R["bias0"], R["jac0"], R["code0"], R["mask1"], R["loc64"], R["homo"], R["feat0"],
                              R["feat1"], R["lo"], a["scale0"], cam, a["L"], a["eps"], w)
        mod.geometric_error(R["R10"], R["t10"], R["bias0"], R["jac0"], R["code0"], a["scale1"] * un, R["mask1"],
                            R["loc64"].to(torch.int32), R["homo"], a["scale0"], cam, a["eps"], a["geo_loss"], a["geo_weight"])

    preps = [prep(p) for p in pairs[:4]]
    for i in range(max(args.warmup, 1)):
        one_pair(*preps[i % len(preps)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.steps):
        one_pair(*preps[s % len(preps)])
    torch.cuda.synchronize()
    t_pair = (time.perf_counter() - t0) / args.steps
    value = 1.0 / (t_pair * npairs)
    print(json.dumps({"impl": args.impl, "metric": METRIC, "value": value, "unit": "LM iters/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": t_pair * npairs * 1e3, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": ("df:: shim over libsage_ba.so" if args.impl == "shim" else "reference CUDA kernels (oracle/_ref)") +
                                             ", pair by pair; each step = 1 ordered pair "
                                             f"(photo+geo linearisation + error evaluation), extrapolated x{npairs}",
                                 "num_samples": args.ref_samples or wl["W"] * wl["H"]},
                      "ms_per_pair": t_pair * 1e3}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu", "shim"])
    ap.add_argument("--small", action="store_true", help="tiny debug workload (not a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tracker", action="store_true", help="skip the configs[1] tracker latency measurement")
    ap.add_argument("--ref-samples", type=int, default=0, help="reference-gpu: sub-sample N points per keyframe (0 = dense)")
    args = ap.parse_args()
    # stdout must carry exactly one JSON line: native libraries (NCCL's version banner, cuSOLVER notes) write to fd 1 behind
    # Python's back, so fd 1 is pointed at stderr and Python's own sys.stdout keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_cpu(args)
    elif args.impl in ("reference-gpu", "shim"):
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""SASS mnemonic census of the hot-path kernels in sage-slam_b200/lib/libsage_ba.so (cuobjdump -sass, sm_100a).

    python profiles/sass_census.py > profiles/r2_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sage-slam_b200", "lib", "libsage_ba.so")
WANT = re.compile(r"photo_kernelILi32ELi32E|geo_kernelILi32E|geo_tc_kernelILi32E|reproj_kernelILi32E|map_match_geom_kernelILi32E|"
                  r"desc_response_kernelILi32E|bs_backward_kernelILi32E|bs_factor_kernelILi32E|assemble_blocks_kernel")
KEEP = re.compile(r"^(FFMA|DFMA|SHFL|LDG|LD\.|LDS|STS|STG|MUFU|BAR|HMMA|DMMA|UTC|LDTM|STTM|SYNCS|UBLKCP|CREDUX|REDUX|REDG|ATOMG|LDGSTS|LDGDEPBAR|"
                  r"NANOSLEEP|FENCE|ELECT)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print("# SASS mnemonic census of sage-slam_b200/lib/libsage_ba.so (cuobjdump -sass, sm_100a): the kernels of the hot path at F = C = 32.\n"
          "# Blackwell-native markers: UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit -> mbarrier), LDTM (tcgen05.ld), UTCATOMSWS (tensor\n"
          "# memory allocation) + SYNCS.* in the geometric lineariser geo_tc_kernel; UBLKCP.S.G (cp.async.bulk: TMA engine) + SYNCS.*\n"
          "# (mbarrier expect_tx / try_wait) in the staged photometric kernels (4th template argument true); DMMA (fp64 tensor cores) +\n"
          "# LDGSTS (cp.async) in the block Cholesky; HMMA.1688.F32.TF32 = the 3xTF32 rank-k updates of the mma.sync factor kernels.\n"
          "# Full listings: cuobjdump -sass -fun <mangled name>.  Regenerate: python profiles/sass_census.py\n")
    name, counts, total = None, None, 0

    def flush():
        if name and WANT.search(name):
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            top = ", ".join(f"{k} {v}" for k, v in sorted(counts.items(), key=lambda t: -t[1]))
            print(f"{demangled}   [{name}]\n  {total} instructions: {top}\n")

    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and name:
            total += 1
            op = m.group(1)
            if KEEP.match(op):
                counts[op] += 1
    flush()


if __name__ == "__main__":
    sys.exit(main())

"""Build libsage_ba.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python sage-slam_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsage_ba.so")
SOURCES = ["photometric.cu", "geometric.cu", "reprojection.cu", "match_geometry.cu", "descriptor.cu", "prep.cu", "api.cu", "problem.cu", "tracker.cu", "blocksolve.cu", "comm.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fvisibility=default"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False, extra_flags=(), lib=None, objdir=None):
    """extra_flags / lib / objdir let tuning scripts build kernel variants side by side (e.g. -DPH_MINB=4)."""
    lib = lib or LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "sage_ba.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src[:-3] + ".o")
        objs.append(o)
        if force or _newer(s, o) or os.path.getmtime(o) < hdr_time:
            jobs.append(["nvcc"] + NVCC_FLAGS + list(extra_flags) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("nvcc failed: " + cmd[-3])
        if verbose and (r.stdout or r.stderr):
            sys.stderr.write(r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    probe_src, probe = os.path.join(CSRC, "tc_probe.cu"), os.path.join(LIBDIR, "tc_probe")
    if lib == LIB and (force or _newer(probe_src, probe) or os.path.getmtime(probe) < hdr_time):
        # known-answer test binary of the tcgen05 conventions (tests/test_gpu_tcgen05.py runs it on the B200)
        run(["nvcc"] + [f for f in NVCC_FLAGS if f not in ("-Xcompiler", "-fPIC", "-fvisibility=default")] + ["-I" + CSRC, "-o", probe, probe_src])
    if jobs or not os.path.exists(lib):
        run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcublas", "-lcusolver", "-lcudart", "-ldl"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

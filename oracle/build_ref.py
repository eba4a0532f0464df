"""Compile the reference's OWN factor kernels into oracle/_ref/ (TEST INFRASTRUCTURE).

The four kernel translation units of the reference
(/root/reference/system/sources/cuda/{photometric,geometric,reprojection,match_geometry}_factor_kernels.cpp)
are compiled where they lie, UNMODIFIED, as CUDA for sm_100a against this image's libtorch,
plus our pybind front (oracle/ref_ext.cpp).  Two shims make that possible without editing
them: ref_shims/sage_ref_compat.h (force-included; re-adds the ::detail::scalar_type
overload that `x.type()` dispatch sites need) and ref_shims/opencv2/opencv.hpp (stub for
PinholeCamera::FromFile).  Nothing is copied into the repo; outputs go to oracle/_ref/
(git-ignored, shipped to the GPU box by gpurun).  The reference has no CPU path, so these
modules can only RUN on a GPU box.

DF_CODE_SIZE / DF_FEAT_SIZE are compile-time in the reference
(system/CMakeLists.txt:41-45), so one module per (CS, FS) is built:
    sage_ref_c{CS}_f{FS}.so
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/system"
OUT = os.path.join(HERE, "_ref")
CONFIGS = [(8, 16), (16, 16), (32, 32)]
SOURCES = ["photometric_factor_kernels.cpp", "geometric_factor_kernels.cpp", "reprojection_factor_kernels.cpp",
           "match_geometry_factor_kernels.cpp"]


def available():
    return os.path.isdir(os.path.join(REF, "sources", "cuda"))


def module_path(cs, fs):
    return os.path.join(OUT, f"sage_ref_c{cs}_f{fs}.so")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-4000:] + "\n")
        raise RuntimeError("reference build failed")


def build(configs=CONFIGS, force=False, jobs=8):
    if not available():
        return False
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda") if hasattr(ce, "include_paths") else ce.include_paths(True):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    inc += ["-I", os.path.join(HERE, "ref_shims"), "-I", os.path.join(REF, "sources", "cuda"),
            "-I", os.path.join(REF, "sources", "common"), "-I", os.path.join(REF, "thirdparty", "eigen")]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    tasks = []
    links = []
    for cs, fs in configs:
        so = module_path(cs, fs)
        if os.path.exists(so) and not force:
            continue
        name = f"sage_ref_c{cs}_f{fs}"
        objs = []
        common = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-x", "cu",
                  "-Xcompiler", "-fPIC", "-w", "--expt-relaxed-constexpr",
                  "-include", os.path.join(HERE, "ref_shims", "sage_ref_compat.h"),
                  f"-DDF_CODE_SIZE={cs}", f"-DDF_FEAT_SIZE={fs}", f"-DTORCH_EXTENSION_NAME={name}",
                  "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))] + inc
        for src in SOURCES:
            obj = os.path.join(OUT, f"{name}_{src[:-4]}.o")
            tasks.append(common + ["-c", os.path.join(REF, "sources", "cuda", src), "-o", obj])
            objs.append(obj)
        obj = os.path.join(OUT, f"{name}_ref_ext.o")
        tasks.append(common + ["-c", os.path.join(HERE, "ref_ext.cpp"), "-o", obj])
        objs.append(obj)
        links.append(["nvcc", "-shared", "-o", so] + objs +
                     ["-L", torch_lib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
                      "-Xlinker", "-rpath", "-Xlinker", torch_lib])
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        list(ex.map(_run, tasks))
    for l in links:
        _run(l)
    for f in os.listdir(OUT):
        if f.endswith(".o") and not f.endswith("_ref_ext.o"):  # the pybind front's object is reused by build_shim()
            os.remove(os.path.join(OUT, f))
    return True


def shim_path(cs, fs):
    return os.path.join(OUT, f"sage_shim_c{cs}_f{fs}.so")


def build_shim(configs=CONFIGS, force=False, jobs=8):
    """The drop-in check of the df:: boundary: integration/df_sage_shim.cpp (the reference's df::*_calculate symbols on top
    of the C ABI) compiled against the reference's own headers and linked with the SAME pybind front (ref_ext.cpp) that
    drives the reference kernels, plus libsage_ba.so -- no reference kernel source is part of this module."""
    if not available():
        return False
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "sage-slam_b200", "lib")
    inc = []
    for p in ce.include_paths("cuda") if hasattr(ce, "include_paths") else ce.include_paths(True):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    inc += ["-I", os.path.join(HERE, "ref_shims"), "-I", os.path.join(REF, "sources", "cuda"),
            "-I", os.path.join(REF, "sources", "common"), "-I", os.path.join(REF, "thirdparty", "eigen"),
            "-I", os.path.join(root, "include")]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    tasks, links = [], []
    for cs, fs in configs:
        so = shim_path(cs, fs)
        shim_src = os.path.join(root, "integration", "df_sage_shim.cpp")
        if os.path.exists(so) and not force and os.path.getmtime(so) > os.path.getmtime(shim_src):
            continue
        name = f"sage_shim_c{cs}_f{fs}"
        common = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-x", "cu",
                  "-Xcompiler", "-fPIC", "-w", "--expt-relaxed-constexpr",
                  "-include", os.path.join(HERE, "ref_shims", "sage_ref_compat.h"),
                  f"-DDF_CODE_SIZE={cs}", f"-DDF_FEAT_SIZE={fs}", f"-DTORCH_EXTENSION_NAME={name}",
                  "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))] + inc
        obj = os.path.join(OUT, f"{name}_df_sage_shim.o")
        tasks.append(common + ["-c", shim_src, "-o", obj])
        # the pybind front is the reference module's object, module name included (load_shim imports the file under that
        # name): build() leaves it behind; compile it here only if it is missing
        ref_name = f"sage_ref_c{cs}_f{fs}"
        front = os.path.join(OUT, f"{ref_name}_ref_ext.o")
        if not os.path.exists(front):
            tasks.append([c if not c.startswith("-DTORCH_EXTENSION_NAME=") else f"-DTORCH_EXTENSION_NAME={ref_name}" for c in common] +
                         ["-c", os.path.join(HERE, "ref_ext.cpp"), "-o", front])
        objs = [obj, front]
        links.append(["nvcc", "-shared", "-o", so] + objs +
                     ["-L", torch_lib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
                      "-L", libdir, "-lsage_ba", "-Xlinker", "-rpath", "-Xlinker", torch_lib,
                      "-Xlinker", "-rpath", "-Xlinker", "/root/repo/sage-slam_b200/lib"])
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        list(ex.map(_run, tasks))
    for l in links:
        _run(l)
    for f in os.listdir(OUT):
        if f.endswith(".o"):
            os.remove(os.path.join(OUT, f))
    return True


def check_mapper_header():
    """Compile-only: include/sage_ba_mapper.hpp (BatchedLocalBA) instantiated on the reference's own df::Map<float>
    (core/mapping/keyframe_map.h) with the vendored Eigen / Sophus and the OpenCV stub.  Raises on failure."""
    if not available():
        return False
    from torch.utils import cpp_extension as ce

    root = os.path.dirname(HERE)
    inc = []
    for p in ce.include_paths("cuda") if hasattr(ce, "include_paths") else ce.include_paths(True):
        inc += ["-isystem", p]
    inc += ["-isystem", "/usr/local/cuda/include", "-I", os.path.join(HERE, "ref_shims"), "-I", os.path.join(REF, "sources", "core", "mapping"),
            "-I", os.path.join(REF, "sources", "common"), "-isystem", os.path.join(REF, "thirdparty", "eigen"),
            "-isystem", os.path.join(REF, "thirdparty", "Sophus"), "-I", os.path.join(root, "include")]
    _run(["g++", "-std=c++17", "-fsyntax-only", "-w"] + inc + [os.path.join(HERE, "check_mapper_header.cpp")])
    return True


def load_shim(cs, fs):
    """Import the shim module (df:: symbols implemented by libsage_ba.so); needs a CUDA device to do anything."""
    import importlib.util
    import torch  # noqa: F401

    path = shim_path(cs, fs)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    name = f"sage_ref_c{cs}_f{fs}"  # the pybind front object is shared with the reference module, so is its init symbol
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(cs, fs):
    """Import the compiled reference module for (CS, FS); needs a CUDA device to do anything."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)

    path = module_path(cs, fs)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    name = f"sage_ref_c{cs}_f{fs}"
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    build_shim(force="--force" in sys.argv)
    check_mapper_header()
    print("reference modules:", sorted(os.listdir(OUT)) if ok else "reference sources not present")

"""The df:: boundary itself: integration/df_sage_shim.cpp defines the reference's `df::*_calculate` symbols on top of the C
ABI.  oracle/build_ref.py compiles it against the reference's own headers and links it with the same pybind front that
drives the reference's kernels (oracle/_ref/sage_shim_c*_f*.so; built where /root/reference exists, shipped to the GPU box).
GPU: the reference's call sequence (oracle/make_golden.run_case) through the shim must reproduce the goldens the
reference's own kernels produced."""
import os
import subprocess

import numpy as np
import pytest

import build_ref
import helpers

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")
SHAPES = {"small_c8_f16": (8, 16), "native_c16_f16": (16, 16), "small_c32_f32": (32, 32)}


def _have_shims():
    return all(os.path.exists(build_ref.shim_path(cs, fs)) for cs, fs in SHAPES.values())


def test_shim_exports_the_reference_symbols_and_binds_the_c_abi():
    if not _have_shims():
        assert not build_ref.available(), "reference present but the shim modules are not built: run __graft_entry__.build()"
        pytest.skip("shim modules are built only where /root/reference exists")
    so = build_ref.shim_path(32, 32)
    defined = subprocess.run(["nm", "-D", "--defined-only", "-C", so], capture_output=True, text=True).stdout
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    for sym in ("df::photometric_jac_error_calculate<32, 32>", "df::photometric_error_calculate<32>",
                "df::geometric_jac_error_calculate<32>", "df::reprojection_jac_error_calculate<32>",
                "df::tracker_photo_jac_error_calculate<32>", "df::tracker_reproj_jac_error_calculate",
                "df::match_geometry_jac_error_calculate<32>", "df::loop_mg_jac_error_calculate"):
        assert sym in defined, sym
    for sym in ("sage_ba_photometric_jac_error", "sage_ba_geometric_jac_error", "sage_ba_reprojection_jac_error",
                "sage_ba_tracker_photo_jac_error", "sage_ba_match_geometry_jac_error", "sage_ba_loop_mg_jac_error",
                "sage_ba_keyframe_create"):
        assert sym in undefined, sym
    assert "photometric_jac_error_calculate_kernel" not in defined  # none of the reference's kernels are inside


@pytest.mark.gpu
@pytest.mark.parametrize("name,far", [(n, f) for n in SHAPES for f in (False, True)])
def test_reference_call_sequence_through_the_shim_matches_goldens(name, far):
    if not _have_shims():
        pytest.skip("shim modules are built only where /root/reference exists (they travel to the GPU box with gpurun)")
    import make_golden

    cs, fs = SHAPES[name]
    mod = build_ref.load_shim(cs, fs)
    out = make_golden.run_case(name, far, mod=mod)
    ref = dict(np.load(os.path.join(GOLDEN, f"{name}{'_far' if far else ''}.npz")))
    np.testing.assert_allclose(out["sig"], ref["sig"], rtol=1e-6)
    checked = 0
    for k, v in ref.items():
        if k == "sig":
            continue
        e = helpers.rel_err(np.asarray(out[k], np.float64).reshape(-1), np.asarray(v, np.float64).reshape(-1))
        assert e <= 1e-4, f"{name} far={far}: {k} rel err {e:.3e}"
        checked += 1
    assert checked >= 40


def _case_tensors(name, dev, make_golden):
    import torch

    kfs = helpers.build_case(name)
    a = helpers.case_args(kfs)
    ta = helpers.tracker_args(kfs, a)
    T = make_golden.T
    R = {k: T(a[k], dev) for k in ("R10", "t10", "R0", "t0", "R1", "t1", "bias0", "code0", "mask1", "homo", "feat0", "feat1", "grad1")}
    R["jac0"] = T(np.ascontiguousarray(a["jac0"].T), dev).t()
    R["loc64"] = T(a["loc1d"], dev, torch.int64)
    R["lo"] = T(a["level_offsets"], dev, torch.int32)
    R["sf"], R["dp"] = T(ta["sfeat0"], dev), T(ta["dpts0"], dev)
    return a, R


@pytest.mark.gpu
def test_shim_tracked_frame_becomes_keyframe_and_cache_evicts_released_frames():
    """The live system's order of calls: a frame is TRACKED first (tracker operators see its feature / gradient pyramids and mask,
    no depth code) and, once it becomes a keyframe, the SAME tensors arrive as frame 1 of the photometric operators
    (core/mapping/mapper.cpp:1274-1275 shares them) -- one cached device copy must serve both, whatever the code size; and
    frames that were released must leave the cache (a long sequence must not grow device memory without bound)."""
    if not _have_shims():
        pytest.skip("shim modules are built only where /root/reference exists (they travel to the GPU box with gpurun)")
    import ctypes
    import gc

    import torch

    import make_golden

    name, (cs, fs) = "native_c16_f16", SHAPES["native_c16_f16"]
    mod = build_ref.load_shim(cs, fs)
    size = ctypes.CDLL(build_ref.shim_path(cs, fs)).sage_shim_cache_size
    size.restype = ctypes.c_long
    dev = torch.device("cuda:0")
    ref = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    a, R = _case_tensors(name, dev, make_golden)
    cam = [float(x) for x in a["cam"]]
    w_cpu = torch.tensor(a["weights"])
    # tracker first (error-only: cached WITHOUT gradients), then with gradients (upgrade in place), then the mapping operator
    gc.collect()  # tensors of earlier tests in this process: their entries are swept by the first miss below
    e_only = mod.tracker_photo_error(R["R10"], R["t10"], R["mask1"], R["dp"], R["homo"], R["sf"], R["feat1"], R["lo"], cam, a["L"], a["eps"],
                                     w_cpu.to(dev))
    n_a = size()  # (the miss above also swept entries of earlier tests whose tensors are gone)
    AtA, Atb, e = mod.tracker_photo_jac_error(R["R10"], R["t10"], R["mask1"], R["dp"], R["homo"], R["sf"], R["feat1"], R["grad1"], R["lo"],
                                              cam, a["L"], a["eps"], w_cpu.to(dev))
    assert helpers.rel_err(AtA.cpu().numpy(), ref["trk_AtA"]) <= 1e-4 and abs(e - ref["trk_err"]) <= 1e-4 * ref["trk_err"]
    assert abs(e_only - ref["trk_err_only"]) <= 1e-4 * ref["trk_err_only"]
    assert size() <= n_a  # still ONE entry for frame 1: upgraded in place with the gradients, not one per operator subset
    n_a = size()
    AtA, Atb, e = mod.photometric_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], R["jac0"], R["code0"],
                                            R["mask1"], R["loc64"], R["homo"], R["feat0"], R["feat1"], R["grad1"], R["lo"], a["scale0"],
                                            cam, a["L"], a["eps"], w_cpu)
    assert helpers.rel_err(AtA.cpu().numpy(), ref["photo_AtA"]) <= 1e-4 and abs(e - ref["photo_err"]) <= 1e-4 * ref["photo_err"]
    assert size() == n_a + 1  # + keyframe 0 (features + depth + samples); frame 1's entry was reused although its C differs
    # a stream of new frames: fresh tensors every time, the previous ones released
    for k in range(6):
        f1 = (R["feat1"] + 0.01 * k).clone()
        g1 = R["grad1"].clone()
        m1 = R["mask1"].clone()
        mod.tracker_photo_jac_error(R["R10"], R["t10"], m1, R["dp"], R["homo"], R["sf"], f1, g1, R["lo"], cam, a["L"], a["eps"], w_cpu.to(dev))
        del f1, g1, m1
        gc.collect()
    assert size() <= n_a + 2, size()  # released frames were evicted on the next miss

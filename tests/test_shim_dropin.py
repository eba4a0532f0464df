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

"""GPU: the TMA-staged photometric kernels of the batched problem (levels >= 1 read from cp.async.bulk windows in shared
memory, tile-major sample order) against the direct-gather kernels and against the CPU oracle, at the BASELINE image size
(320x256, F = C = 32, L = 4): dense, endoscope mask, sub-sampled (windows clipped -> mixed shared / global taps), and with a
strong in-plane rotation between the two frames (the windows of a 32x4 tile no longer fit the budget)."""
import os

import numpy as np
import pytest

import helpers
import oracle as O
import sage_slam_b200 as sage
from sage_slam_b200 import local_ba

PRM = dict(W=320, H=256, L=4, F=32, C=32, seed=31)


def _scene(mask="full", num_samples=None, roll_deg=0.0, num_kf=3):
    kfs = sage.synthetic.make_scene(num_kf=num_kf, mask=mask, num_samples=num_samples, **PRM)
    rng = np.random.default_rng(5)
    for k in kfs:
        k.code = (0.2 * rng.standard_normal(PRM["C"])).astype(np.float32)
        k.dpt_scale = float(np.float32(1.0 + 0.05 * rng.standard_normal()))
    if roll_deg:
        a = np.deg2rad(roll_deg)
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
        R, t = kfs[1].pose_wk
        kfs[1].pose_wk = ((R @ Rz).astype(np.float32), t)
    return kfs


def _problem(ctx, dk, kfs, pairs, staged):
    os.environ["SAGE_BA_STAGED"] = "1" if staged else "0"
    try:
        ba = sage.LocalBA(ctx, dk)
    finally:
        os.environ.pop("SAGE_BA_STAGED", None)
    for i, j in pairs:
        ba.add_photometric(i, j, helpers.PHOTO_WEIGHTS[:PRM["L"]])
    ba.set_state([k.pose_wk for k in kfs], np.stack([k.code for k in kfs]), [k.dpt_scale for k in kfs], helpers.EPS)
    return ba


def _factor_outputs(ba, npairs, C):
    ba.linearize()
    buf = ba.factor_buffer()
    offs, dims, _ = local_ba.factor_layout(["photo"] * npairs, C)
    out = []
    for off, D in zip(offs, dims):
        out.append((buf[off:off + D * D].reshape(D, D).copy(), buf[off + D * D:off + D * D + D].copy(), float(buf[off + D * D + D]),
                    float(buf[off + D * D + D + 1])))
    ba.evaluate(candidate=False)
    return out, ba._buffer_view("cost")[2].cpu().numpy().copy()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["dense", "ellipse", "subsampled", "roll25"])
def test_staged_kernels_match_direct_kernels_and_oracle(sage_ctx, variant):
    kfs = _scene(mask="ellipse" if variant == "ellipse" else "full", num_samples=3072 if variant == "subsampled" else None,
                 roll_deg=25.0 if variant == "roll25" else 0.0)
    pairs = [(0, 1), (1, 0), (1, 2), (2, 0)]
    C = PRM["C"]
    dk = [sage.DeviceKeyframe(sage_ctx, k) for k in kfs]
    direct, dcost = _factor_outputs(_problem(sage_ctx, dk, kfs, pairs, False), len(pairs), C)
    staged, scost = _factor_outputs(_problem(sage_ctx, dk, kfs, pairs, True), len(pairs), C)
    for f, ((A0, b0, e0, n0), (A1, b1, e1, n1)) in enumerate(zip(direct, staged)):
        assert n0 == n1 and n0 > 0, (variant, f, n0, n1)
        assert abs(e1 - e0) <= 2e-6 * abs(e0), (variant, f, e0, e1)
        helpers.assert_blocks_close("photo", A1, b1, A0, b0, C, 2e-5, f"{variant} factor {f} staged vs direct")
    # error-only kernels: [error | inliers] per factor
    np.testing.assert_array_equal(dcost[1::2], scost[1::2])
    assert np.abs(dcost[0::2] - scost[0::2]).max() <= 2e-6 * np.abs(dcost[0::2]).max()
    # and against the CPU oracle on the first pair (fp32 rows, fp64 sums)
    a = helpers.case_args(kfs, *pairs[0])
    Ao, bo, eo, no = O.photometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                             a["mask1"], a["loc1d"], a["homo"], a["feat0"], a["feat1"], a["grad1"], a["level_offsets"],
                                             a["scale0"], a["cams"], a["eps"], a["weights"])
    A1, b1, e1, n1 = staged[0]
    assert n1 == no and abs(e1 - eo) <= 1e-4 * eo
    helpers.assert_blocks_close("photo", A1, b1, Ao, bo, C, 1e-4, f"{variant} staged vs oracle")
    for x in dk:
        x.close()

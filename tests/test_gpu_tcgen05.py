"""GPU: the tcgen05 geometric lineariser (csrc/geometric.cu: geo_tc_kernel -- tcgen05.mma kind::tf32, accumulators in tensor
memory, rows staged K-major in shared memory) against the mma.sync lineariser, the goldens of the reference's own kernels and
the CPU oracle; and the known-answer probe that pins the tcgen05 conventions the kernel is built on."""
import os
import subprocess

import numpy as np
import pytest

import helpers
import sage_slam_b200 as sage
from sage_slam_b200 import capi, local_ba, ops

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")
PROBE = os.path.join(helpers.ROOT, "sage-slam_b200", "lib", "tc_probe")


@pytest.fixture
def tcgen05():
    """Switch the process-wide lineariser choice for one test and restore it."""
    lib = capi.load()
    prev = lib.sage_ba_set_geometric_tcgen05(-1)

    def use(on):
        lib.sage_ba_set_geometric_tcgen05(1 if on else 0)

    yield use
    lib.sage_ba_set_geometric_tcgen05(prev)


@pytest.mark.gpu
@pytest.mark.parametrize("test,variant,reps", [(0, 0, 1), (2, 0, 1), (2, 2, 5), (3, 0, 1), (3, 2, 5)])
def test_probe_pins_the_tcgen05_conventions(test, variant, reps):
    """tensor-memory st/ld round trip; K-major no-swizzle descriptors with the natural (256 / 128) and the padded (272 / 144)
    strides geo_tc_kernel uses; accumulation into a pre-filled accumulator.  Exact integer answers (M=128, N=160, K=128)."""
    assert os.path.exists(PROBE), "run __graft_entry__.build()"
    r = subprocess.run([PROBE, str(test), str(variant), str(reps)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0 of 20480" in r.stdout, r.stdout


def _geo(ctx, kfs, jac=True):
    a = helpers.case_args(kfs)
    d0, d1 = sage.DeviceKeyframe(ctx, kfs[0]), sage.DeviceKeyframe(ctx, kfs[1])
    try:
        return ops.geometric_jac_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"], a["code1"],
                                                 a["scale0"], a["scale1"], a["eps"], a["geo_loss"], a["geo_weight"])
    finally:
        d0.close()
        d1.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,far", [(n, f) for n, f in helpers.GOLDEN_RUNS if helpers.CASES[n]["C"] == 32])
def test_tcgen05_lineariser_matches_mma_sync_and_the_reference_goldens(sage_ctx, tcgen05, name, far):
    kfs = helpers.build_case(name, far)
    C = helpers.CASES[name]["C"]
    tcgen05(False)
    A0, b0, e0, n0 = _geo(sage_ctx, kfs)
    tcgen05(True)
    before = sage_ctx.launch_count
    A1, b1, e1, n1 = _geo(sage_ctx, kfs)
    assert sage_ctx.launch_count > before
    assert n1 == n0 and abs(e1 - e0) <= 2e-6 * abs(e0), (e0, e1, n0, n1)
    helpers.assert_blocks_close("geo", A1, b1, A0, b0, C, 2e-5, f"{name} far={far}: tcgen05 vs mma.sync")
    ref = dict(np.load(os.path.join(GOLDEN, f"{name}{'_far' if far else ''}.npz")))
    assert abs(e1 - float(ref["geo_err"])) <= 1e-4 * abs(float(ref["geo_err"]))
    helpers.assert_blocks_close("geo", A1, b1, ref["geo_AtA"], ref["geo_Atb"].reshape(-1), C, 1e-4, f"{name} far={far}: tcgen05 vs reference golden")


@pytest.mark.gpu
@pytest.mark.parametrize("num_samples", [None, 3000])
def test_tcgen05_lineariser_in_the_batched_problem(sage_ctx, tcgen05, num_samples):
    """All geometric factors of a problem in one launch (ragged tail: 3000 samples is not a multiple of the 128-sample round)."""
    prm = dict(W=160, H=128, L=3, F=16, C=32, seed=41)
    kfs = sage.synthetic.make_scene(num_kf=4, mask="ellipse", num_samples=num_samples, **prm)
    rng = np.random.default_rng(6)
    for k in kfs:
        k.code = (0.2 * rng.standard_normal(prm["C"])).astype(np.float32)
        k.dpt_scale = float(np.float32(1.0 + 0.05 * rng.standard_normal()))
    pairs = [(0, 1), (1, 0), (1, 2), (2, 3), (3, 0)]
    dk = [sage.DeviceKeyframe(sage_ctx, k) for k in kfs]
    bufs = []
    for on in (False, True):
        tcgen05(on)
        ba = sage.LocalBA(sage_ctx, dk)
        for i, j in pairs:
            ba.add_geometric(i, j, 0.1, 1.0)
        ba.set_state([k.pose_wk for k in kfs], np.stack([k.code for k in kfs]), [k.dpt_scale for k in kfs], helpers.EPS)
        tcgen05(not on)  # flipping the process-wide switch after creation must not touch this problem (its buffers are sized)
        ba.linearize()
        bufs.append(ba.factor_buffer().copy())
        ba.close()
    offs, dims, _ = local_ba.factor_layout(["geo"] * len(pairs), prm["C"])
    for f, (off, D) in enumerate(zip(offs, dims)):
        A0, A1 = (b[off:off + D * D].reshape(D, D) for b in bufs)
        g0, g1 = (b[off + D * D:off + D * D + D] for b in bufs)
        assert bufs[0][off + D * D + D + 1] == bufs[1][off + D * D + D + 1] > 0  # inliers
        assert abs(bufs[1][off + D * D + D] - bufs[0][off + D * D + D]) <= 2e-6 * abs(bufs[0][off + D * D + D])
        helpers.assert_blocks_close("geo", A1, g1, A0, g0, prm["C"], 2e-5, f"factor {f}: tcgen05 vs mma.sync")
    for x in dk:
        x.close()

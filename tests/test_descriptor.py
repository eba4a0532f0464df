"""Row f3: dense descriptor cycle-matching.  CPU: the numpy oracle against the torch replay of the reference's
expression chain (tests/golden/desc_*.npz, oracle/make_golden_desc.py) and against properties.  GPU: the sm_100a
kernels through the C ABI against the oracle and the goldens -- indices must be EXACT."""
import os

import numpy as np
import pytest

import desc_case
import helpers
import oracle

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")
KEYS = ("raw_matched_locations_1d_1", "cyc_matched_locations_1d_0", "inlier_within_keypoint_indexes", "matched_locations_1d_1",
        "matched_locations_2d_1")


def _golden(device):
    fn = os.path.join(GOLDEN, f"desc_{device}.npz")
    assert os.path.exists(fn), "golden fixture missing: run oracle/make_golden_desc.py"
    return dict(np.load(fn))


@pytest.mark.parametrize("device", ["cpu", "cuda"])
@pytest.mark.parametrize("name", list(desc_case.CASES))
def test_oracle_reproduces_reference_chain(name, device):
    g = _golden(device)
    c = desc_case.build(name)
    np.testing.assert_allclose(desc_case.signature(c), g[f"{name}/sig"], rtol=1e-12)
    o = oracle.cycle_match(c["desc0"], c["desc1"], c["kp"], c["thresh"])
    for k in KEYS:
        np.testing.assert_array_equal(np.asarray(o[k]), g[f"{name}/{k}"], err_msg=f"{name}: {k}")
    n_in = len(o["inlier_within_keypoint_indexes"])
    assert 0 < n_in <= c["K"]


def test_oracle_properties():
    c = desc_case.build("d16_small")
    # a map matched against itself: every keypoint finds itself, both ways
    o = oracle.cycle_match(c["desc0"], c["desc0"], c["kp"], 0.0)
    np.testing.assert_array_equal(o["raw_matched_locations_1d_1"], c["kp"])
    np.testing.assert_array_equal(o["cyc_matched_locations_1d_0"], c["kp"])
    assert len(o["inlier_within_keypoint_indexes"]) == c["K"]
    # a pure cyclic shift without noise: the match is the shifted pixel
    H, W = c["H"], c["W"]
    d1 = np.roll(c["desc0"], (2, 3), axis=(1, 2))
    o = oracle.cycle_match(c["desc0"], d1, c["kp"], 0.0)
    y, x = c["kp"] // W, c["kp"] % W
    np.testing.assert_array_equal(o["raw_matched_locations_1d_1"], ((y + 2) % H) * W + (x + 3) % W)
    # constant maps: every response ties, argmax keeps the first pixel, nothing is cycle-consistent except pixel 0
    z = np.ones_like(c["desc0"])
    o = oracle.cycle_match(z, z, np.array([0, 5, 77]), 1.0)
    np.testing.assert_array_equal(o["raw_matched_locations_1d_1"], [0, 0, 0])
    np.testing.assert_array_equal(o["inlier_within_keypoint_indexes"], [0])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(desc_case.CASES))
def test_kernels_match_oracle_and_goldens_exactly(sage_ctx, name):
    from sage_slam_b200 import ops

    c = desc_case.build(name)
    mine = ops.cycle_feature_matching(sage_ctx, c["desc0"], c["desc1"], c["kp"], c["thresh"])
    o = oracle.cycle_match(c["desc0"], c["desc1"], c["kp"], c["thresh"])
    for k in KEYS:
        np.testing.assert_array_equal(np.asarray(mine[k]), np.asarray(o[k]), err_msg=f"{name}: {k} vs oracle")
    for device in ("cpu", "cuda"):
        g = _golden(device)
        for k in KEYS:
            np.testing.assert_array_equal(np.asarray(mine[k]), g[f"{name}/{k}"], err_msg=f"{name}: {k} vs golden {device}")


@pytest.mark.gpu
def test_ties_errors_and_device_pointers(sage_ctx):
    import torch

    from sage_slam_b200 import ops

    c = desc_case.build("d16_small")
    z = np.ones_like(c["desc0"])
    r = ops.cycle_feature_matching(sage_ctx, z, z, np.array([0, 5, 77]), 1.0)
    np.testing.assert_array_equal(r["raw_matched_locations_1d_1"], [0, 0, 0])
    np.testing.assert_array_equal(r["inlier_within_keypoint_indexes"], [0])
    with pytest.raises(ops.SageError):
        ops.cycle_feature_matching(sage_ctx, c["desc0"], c["desc1"], np.array([c["H"] * c["W"]]), 2.0)
    with pytest.raises(ops.SageError):
        ops.cycle_feature_matching(sage_ctx, c["desc0"][:5], c["desc1"][:5], c["kp"], 2.0)  # 5 channels unsupported
    # maps already on the device (as Frame::feat_desc is in the live system)
    t0, t1 = torch.from_numpy(c["desc0"]).cuda(), torch.from_numpy(c["desc1"]).cuda()
    torch.cuda.synchronize()
    r = ops.cycle_feature_matching(sage_ctx, None, None, c["kp"], c["thresh"], device_ptrs=(t0.data_ptr(), t1.data_ptr(), c["C"], c["H"], c["W"]))
    o = oracle.cycle_match(c["desc0"], c["desc1"], c["kp"], c["thresh"])
    for k in KEYS:
        np.testing.assert_array_equal(np.asarray(r[k]), np.asarray(o[k]))


@pytest.mark.gpu
def test_full_size_round_trip(sage_ctx):
    """320x256x32 descriptors, K = 512 (desc_num_keypoints, slam_run.flags:97): too big for the numpy oracle in seconds, so
    check the size-independent property: matching a map against a cyclic shift of itself returns the shifted pixel and the
    return pass comes back to the keypoint."""
    from sage_slam_b200 import ops

    rng = np.random.default_rng(5)
    C, H, W, K = 32, 256, 320, 512
    d0 = rng.standard_normal((C, H, W)).astype(np.float32)
    d1 = np.roll(d0, (5, -7), axis=(1, 2))
    kp = rng.permutation(H * W)[:K].astype(np.int64)
    r = ops.cycle_feature_matching(sage_ctx, d0, d1, kp, 0.0, timing=True)
    y, x = kp // W, kp % W
    np.testing.assert_array_equal(r["raw_matched_locations_1d_1"], ((y + 5) % H) * W + (x - 7) % W)
    np.testing.assert_array_equal(r["cyc_matched_locations_1d_0"], kp)
    assert len(r["inlier_within_keypoint_indexes"]) == K
    assert r["kernel_ms"] > 0


# ---------------------------------------------------------------------------------------------------------------------------
# The reference's OWN statements (core/gtsam/reprojection_factor.cpp:42-89, extracted at build time, libtorch CPU):
# tests/golden/match_pins.npz, oracle/match_pins.cpp, oracle/make_golden_match.py
@pytest.mark.parametrize("name", list(desc_case.CASES))
def test_keypoint_draw_and_cycle_matching_match_the_references_own_statements(name):
    import make_golden_match as M
    import make_golden_desc as D
    import sage_slam_b200 as sage
    import torch

    pins = np.load(os.path.join(GOLDEN, "match_pins.npz"))
    c = desc_case.build(name)
    loc = M.valid_locations(c)
    kf_id, fr_id = M.IDS[name]
    # 1. the keypoint draw: std::shuffle(iota(N), mt19937(kf id * frame id)), first K (the reference image links libstdc++ of GCC 9;
    #    this container's g++ 13 built the harness, and frames.std_shuffle restates both)
    drawn = pins[name + "/keypoint_indexes"]
    np.testing.assert_array_equal(sage.frames.std_shuffle(len(loc), kf_id * fr_id, libstdcxx="13")[:c["K"]], drawn)
    kp = loc[drawn]
    # 2. the cycle matching on those keypoints: the torch replay that generated desc_*.npz, and the oracle's restatement
    with torch.no_grad():
        r = D.reference_chain(torch.from_numpy(c["desc0"][None]), torch.from_numpy(c["desc1"][None]), torch.from_numpy(kp), c["W"], c["H"],
                              c["thresh"])
    o = oracle.cycle_match(c["desc0"], c["desc1"], kp, c["thresh"])
    for got in (r, o):
        np.testing.assert_array_equal(np.asarray(got["raw_matched_locations_1d_1"]), pins[name + "/raw_matched_locations_1d_1"])
        np.testing.assert_array_equal(np.asarray(got["cyc_matched_locations_1d_0"]), pins[name + "/cyc_matched_locations_1d_0"])
        inl = np.asarray(got["inlier_within_keypoint_indexes"])
        np.testing.assert_array_equal(drawn[inl], pins[name + "/matched_keypoint_indexes"])

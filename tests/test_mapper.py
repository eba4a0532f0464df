"""Row f4: MappingStep / UpdateMap adapter.  GPU: a 4-keyframe map is built through the Mapper-style calls
(init_one_frame, enqueue_keyframe), optimised by one mapping_step and written back; the written-back depth maps must equal
UpdateDepth of the written-back code/scale (oracle.update_depth), the gauge keyframe must not move, the cost must fall."""
import numpy as np
import pytest

import oracle
import sage_slam_b200 as sage
from sage_slam_b200.mapper import BatchedMapper, MapperOptions


def _scene():
    kfs = sage.synthetic.make_scene(num_kf=4, W=64, H=48, L=3, F=16, C=8, seed=21, with_desc=True)
    # what tracking + CorrectDepthScale hand to the mapper: depths (and therefore translations) in units of KF0's median depth
    valid = kfs[0].dpt_map.reshape(-1)[kfs[0].sampled_locations_1d]
    m = np.float32(np.sort(valid)[(len(valid) - 1) // 2])
    for kf in kfs:
        kf.dpt_scale = float(np.float32(1.0) / m)
        kf.pose_wk = (kf.pose_wk[0], (kf.pose_wk[1] / m).astype(np.float32))
        kf.pose_wk_true = (kf.pose_wk_true[0], (kf.pose_wk_true[1] / m).astype(np.float32))
    return kfs


def _rot_err_deg(kfs):
    """largest rotation error against the ground truth (translations are only defined up to the free global depth scale)"""
    out = 0.0
    for k in kfs[1:]:
        c = (np.trace(k.pose_wk[0].astype(np.float64).T @ k.pose_wk_true[0].astype(np.float64)) - 1.0) / 2.0
        out = max(out, float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0)))))
    return out


def test_keyframe_dpt_map_follows_update_depth():
    kfs = sage.synthetic.make_scene(num_kf=1, W=32, H=24, L=2, F=16, C=8, seed=3)
    kf = kfs[0]
    kf.code = np.linspace(-0.2, 0.2, 8).astype(np.float32)
    kf.dpt_scale = 1.3
    np.testing.assert_allclose(kf.dpt_map.reshape(-1), oracle.update_depth(kf.dpt_map_bias, kf.dpt_jac_code, kf.code, kf.dpt_scale),
                               rtol=1e-6)
    kf.dpt_map_stored = np.zeros((24, 32), np.float32)
    assert kf.dpt_map is kf.dpt_map_stored  # UpdateMap's stored map wins, as Frame::dpt_map does in the reference
    o = MapperOptions()
    assert o.photo_factor_weights == (10.0, 9.0, 8.0, 7.0) and o.desc_num_keypoints == 512  # configs/slam_run.flags:96-106


@pytest.mark.gpu
def test_mapping_step_and_update_map(sage_ctx):
    kfs = _scene()
    mp = BatchedMapper(sage_ctx, MapperOptions(desc_num_keypoints=128, factor_iters=8))
    mp.init_one_frame(kfs[0])
    assert abs(kfs[0].dpt_scale * float(np.median(kfs[0].dpt_map_bias)) - 1.0) < 0.2
    for kf in kfs[1:]:
        mp.enqueue_keyframe(kf, kf.temporal_connections)
    kinds = [f[0] for f in mp._factors]
    n_links = sum(len(k.temporal_connections) for k in kfs[1:])
    assert kinds.count("photo") == 2 * n_links and kinds.count("geo") == 2 * n_links
    assert kinds.count("reproj") == 2 * n_links, mp.match_stats
    assert all(n > 0.5 * K for n, K in mp.match_stats.values()), mp.match_stats  # consistent synthetic views: most keypoints match
    R0, t0 = kfs[0].pose_wk[0].copy(), kfs[0].pose_wk[1].copy()
    rep = mp.mapping_step()
    assert rep["final_cost"] < 0.5 * rep["initial_cost"], rep
    assert rep["accepted"] >= 1
    np.testing.assert_array_equal(kfs[0].pose_wk[0], R0)  # gauge keyframe: pose held
    np.testing.assert_array_equal(kfs[0].pose_wk[1], t0)
    assert _rot_err_deg(kfs) < 1.0
    for kf in kfs:
        want = oracle.update_depth(kf.dpt_map_bias, kf.dpt_jac_code, kf.code, kf.dpt_scale)
        np.testing.assert_allclose(kf.dpt_map.reshape(-1), want, rtol=2e-6, atol=1e-7)
        assert kf.dpt_map.shape == kf.video_mask.shape
    # a second step starts from the written-back state and must not get worse
    rep2 = mp.mapping_step(iters=3)
    assert rep2["final_cost"] <= rep["final_cost"] * (1 + 1e-6)
    mp.close()

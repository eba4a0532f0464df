"""Shared construction of parity-test cases (used by tests/, oracle/make_golden.py, bench.py's baselines).

A *case* is a seeded synthetic two-keyframe scene plus a perturbed state; `case_args` flattens it into the
reference-layout arrays every implementation (reference CUDA kernels, CPU oracle, sm_100a kernels) consumes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import sage_slam_b200 as sage  # noqa: E402

F32 = np.float32

# name -> scene parameters.  (C, F) pairs match the compiled reference modules (8,16) (16,16) (32,32).
CASES = {
    "small_c8_f16": dict(W=64, H=48, L=3, F=16, C=8, num_samples=None, mask="full", seed=11),
    "native_c16_f16": dict(W=80, H=64, L=4, F=16, C=16, num_samples=3072, mask="ellipse", seed=12),
    "small_c32_f32": dict(W=64, H=48, L=4, F=32, C=32, num_samples=None, mask="full", seed=13),
    # BASELINE.json sizes: configs[0] (128x96, F16, C8, L=4, dense) and the bench shape of configs[1..4] (320x256, F=C=32, L=4,
    # dense N=81920), the latter also under the endoscope (ellipse) mask
    "cfg0_c8_f16": dict(W=128, H=96, L=4, F=16, C=8, num_samples=None, mask="full", seed=14),
    "bench_c32_f32": dict(W=320, H=256, L=4, F=32, C=32, num_samples=None, mask="full", seed=15),
    "bench_ell_c32_f32": dict(W=320, H=256, L=4, F=32, C=32, num_samples=None, mask="ellipse", seed=16),
}
# (case, far) combinations with a golden fixture; far = KF1 moved away so that nothing overlaps
GOLDEN_RUNS = [(n, far) for n in CASES for far in (False, True) if not (far and n in ("cfg0_c8_f16", "bench_ell_c32_f32"))]
PHOTO_WEIGHTS = [10.0, 9.0, 8.0, 7.0]
EPS = 1e-4


def build_case(name, far=False):
    """Two keyframes + state.  far=True moves KF1 so that nothing overlaps (zero-inlier fallback)."""
    prm = CASES[name]
    kfs = sage.synthetic.make_scene(num_kf=2, **prm)
    rng = np.random.default_rng(prm["seed"] + 100)
    C = prm["C"]
    for k in kfs:
        k.code = (0.3 * rng.standard_normal(C)).astype(F32)
        k.dpt_scale = float(F32(1.0 + 0.1 * rng.standard_normal()))
    if far:
        R, t = kfs[1].pose_wk
        kfs[1].pose_wk = (R, (t + np.array([50.0, 0, 0])).astype(F32))
    return kfs


def rel_pose_f32(p0, p1):
    R0, t0 = p0
    R1, t1 = p1
    R10 = (R1.T.astype(F32) @ R0.astype(F32)).astype(F32)
    t10 = (R1.T.astype(F32) @ (t0 - t1).astype(F32)).astype(F32)
    return R10, t10


def case_args(kfs, i=0, j=1):
    """Reference-layout argument dict for the ordered pair i -> j."""
    a, b = kfs[i], kfs[j]
    R10, t10 = rel_pose_f32(a.pose_wk, b.pose_wk)
    cams = a.camera_pyramid
    L = len(cams)
    H, W = a.video_mask.shape
    C = a.dpt_jac_code.shape[1]
    D1u = (b.dpt_map_bias + b.dpt_jac_code @ b.code).reshape(H, W).astype(F32)  # unscaled depth of KF1
    gx, gy = sage.frames.spatial_grad(D1u[None])
    s1 = F32(b.dpt_scale)
    return dict(
        R10=R10, t10=t10, R0=a.pose_wk[0], t0=a.pose_wk[1], R1=b.pose_wk[0], t1=b.pose_wk[1],
        bias0=a.dpt_map_bias, jac0=a.dpt_jac_code, bias1=b.dpt_map_bias, jac1=b.dpt_jac_code, code0=a.code, code1=b.code, scale0=float(a.dpt_scale),
        scale1=float(b.dpt_scale), mask1=b.video_mask, loc1d=a.sampled_locations_1d, homo=a.sampled_locations_homo,
        feat0=a.feat_map_pyramid, feat1=b.feat_map_pyramid, grad1=b.feat_map_grad_pyramid,
        level_offsets=a.level_offsets, cams=cams, eps=EPS, weights=np.array(PHOTO_WEIGHTS[:L], F32),
        # what GeometricFactor::ComputeJacobianAndError hands over (geometric_factor.cpp:317-320,340-342)
        dpt1=(s1 * D1u).astype(F32), dgrad1=(s1 * np.stack([gx[0], gy[0]])).astype(F32),
        basis1=np.ascontiguousarray(b.dpt_jac_code.reshape(H, W, C)), cam=cams[0],
        geo_loss=float(0.03 * np.mean(a.dpt_map_bias.astype(np.float64) ** 2)), geo_weight=0.1,
        rep_loss=float(0.03 * W * W), rep_weight=0.1, H=H, W=W, C=C, L=L, F=a.feat_map_pyramid.shape[0])


def tracker_args(kfs, args):
    """The tracker's pre-sampled tensors (camera_tracker.cpp:1086-1123) via torch grid_sample on the CPU."""
    import torch
    import torch.nn.functional as Fn

    a = kfs[0]
    H, W, L = args["H"], args["W"], args["L"]
    loc = torch.from_numpy(a.sampled_locations_1d)
    x = torch.fmod(loc.to(torch.float32), W)
    y = torch.floor(loc.to(torch.float32) / W)
    grid = torch.stack([(x + 0.5) * (2.0 / W) - 1.0, (y + 0.5) * (2.0 / H) - 1.0], 1).reshape(1, 1, -1, 2)
    feats = []
    pyr = torch.from_numpy(a.feat_map_pyramid)
    Fc = pyr.shape[0]
    for l in range(L):
        w, h = int(args["cams"][l][4]), int(args["cams"][l][5])
        off = int(a.level_offsets[l])
        m = pyr[:, off:off + w * h].reshape(1, Fc, h, w)
        s = Fn.grid_sample(m, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        feats.append(s.reshape(Fc, -1).permute(1, 0))
    sfeat0 = torch.stack(feats, 0).contiguous().numpy()
    dpts0 = a.dpt_map.reshape(-1)[a.sampled_locations_1d].astype(F32)
    return dict(sfeat0=sfeat0, dpts0=dpts0)


def match_args(kfs, M=200):
    loc, homo, uv = sage.synthetic.make_matches(kfs[0], kfs[1], M=M)
    dpts = kfs[0].dpt_map.reshape(-1)[loc].astype(F32)
    # 3-D form of the same matches for the match-geometry term: ray + depth of the matched pixel in frame 1
    cam = kfs[1].camera_pyramid[0]
    H, W = kfs[1].video_mask.shape
    homo1 = np.stack([(uv[:, 0] - cam[2]) / cam[0], (uv[:, 1] - cam[3]) / cam[1], np.ones(len(uv), F32)], 1).astype(F32)
    px = np.clip(np.round(uv[:, 0]).astype(int), 0, W - 1)
    py = np.clip(np.round(uv[:, 1]).astype(int), 0, H - 1)
    dpts1 = kfs[1].dpt_map[py, px].astype(F32)
    mg_loss = float(0.03 * np.mean(kfs[0].dpt_map_bias.astype(np.float64) ** 2))
    # mapping / loop-closure forms: 1-D location of the matched pixel in frame 1 and the UNSCALED depths of both keypoints
    loc1 = (py * W + px).astype(np.int32)
    mud0 = (kfs[0].dpt_map_bias[loc] + kfs[0].dpt_jac_code[loc] @ kfs[0].code).astype(F32)
    mud1 = (kfs[1].dpt_map_bias[loc1] + kfs[1].dpt_jac_code[loc1] @ kfs[1].code).astype(F32)
    return dict(mloc=loc, mhomo=homo, m2d=uv, mdpts=dpts, mhomo1=homo1, mdpts1=dpts1, mg_loss=mg_loss, mg_weight=0.2, mloc1=loc1,
                mud0=mud0, mud1=mud1)


MG_LOSSES = ("fair", "L2", "huber", "unbiased")


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def system_blocks(key, D, C):
    """Variable blocks of a factor's normal equations, in the reference's column order (name, start, stop)."""
    if key in ("photo", "rep"):  # [pose0 6 | pose1 6 | code0 C | scale0]   photometric_factor_kernels.cpp:241-364
        b = [("pose0", 0, 6), ("pose1", 6, 12), ("code0", 12, 12 + C), ("scale0", 12 + C, 13 + C)]
    elif key == "geo" or key.startswith("mmg_"):  # [pose0 | pose1 | code0 | code1 | scale0 | scale1]   geometric_factor_kernels.cpp:671-696
        b = [("pose0", 0, 6), ("pose1", 6, 12), ("code0", 12, 12 + C), ("code1", 12 + C, 12 + 2 * C), ("scale0", 12 + 2 * C, 13 + 2 * C),
             ("scale1", 13 + 2 * C, 14 + 2 * C)]
    elif key == "lmg":  # loop closure: [pose0 | pose1 | scale0 | scale1]
        b = [("pose0", 0, 6), ("pose1", 6, 12), ("scale0", 12, 13), ("scale1", 13, 14)]
    else:  # tracker forms: [pose 6 (| scale)]
        b = [("pose", 0, 6)] + ([("scale", 6, 7)] if D == 7 else [])
    assert b[-1][2] == D, (key, D, C)
    return b


def block_errors(key, A, b, A_ref, b_ref, C):
    """Per-block error of a factor's (AtA, Atb) against a reference, in the Jacobi-scaled system: every variable is scaled to
    unit diagonal (A~ = S A S, b~ = S b, S = diag(A_ref)^-1/2), so that pose-pose entries (~ f^2/z^2) cannot hide errors in the
    code / scale blocks -- |A~_ij| <= 1 for every entry of a Gram matrix, and an entry's rounding error scales with the norms
    of its two columns.  Returns {(rowblock, colblock): max |dA~|} and {block: max |db~| / max |b~_ref|}."""
    A, A_ref = np.asarray(A, np.float64), np.asarray(A_ref, np.float64)
    b, b_ref = np.asarray(b, np.float64).reshape(-1), np.asarray(b_ref, np.float64).reshape(-1)
    D = A_ref.shape[0]
    d = np.diag(A_ref).copy()
    live = d > 0
    s = np.where(live, 1.0 / np.sqrt(np.where(live, d, 1.0)), 1.0)
    dA = np.abs(A - A_ref) * s[:, None] * s[None, :]
    bs = np.abs(b_ref * s).max()
    db = np.abs(b - b_ref) * s / max(bs, 1e-30)
    blocks = system_blocks(key, D, C)
    ea = {(r[0], c[0]): float(dA[r[1]:r[2], c[1]:c[2]].max()) for r in blocks for c in blocks}
    eb = {r[0]: float(db[r[1]:r[2]].max()) for r in blocks}
    return ea, eb


def assert_blocks_close(key, A, b, A_ref, b_ref, C, tol, label):
    if not np.any(np.asarray(A_ref)):  # zero-overlap fallback: all-zero system
        assert not np.any(np.asarray(A)) and not np.any(np.asarray(b)), f"{label}: {key} must be an all-zero system"
        return 0.0
    ea, eb = block_errors(key, A, b, A_ref, b_ref, C)
    worst = max(list(ea.values()) + list(eb.values()))
    bad = {k: v for k, v in list(ea.items()) + list(eb.items()) if v > tol}
    assert not bad, f"{label}: {key} blocks over {tol:g} in the Jacobi-scaled system: {bad}"
    return worst

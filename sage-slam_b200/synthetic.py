"""Deterministic procedural scenes for tests / bench / smoke.

There is no dataset access, so every input is synthetic but GEOMETRICALLY CONSISTENT: a
textured height-field surface z = g(x, y) in world coordinates is viewed by K cameras on a
smooth trajectory; depth per pixel comes from ray/surface intersection and feature channel k
is sin(omega_k . X + phi_k) evaluated at the 3-D surface point X, so warped features agree
between views and LM converges.  Pyramids, gradients, sample points and the depth basis are
built exactly as the reference builds them (frames.py).
"""
import os

import numpy as np

from .frames import (F32, Keyframe, camera_pyramid, gaussian_pyramid_with_grad, level_offsets, mask_pyramid,
                     valid_locations)


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _surface(x, y):
    return 2.0 + 0.15 * x - 0.10 * y + 0.12 * np.sin(1.7 * x + 0.3) * np.cos(1.3 * y - 0.2)


def _smooth_field(rng, H, W, sigma):
    """Gaussian-blurred N(0,1) field, unit std; separable FFT-free blur via cumulative box filters."""
    f = rng.standard_normal((H, W))
    r = max(1, int(sigma))
    for _ in range(3):  # 3 box passes ~ Gaussian
        p = np.pad(f, ((r, r), (r, r)), mode="reflect")
        c = np.cumsum(np.cumsum(p, 0), 1)
        c = np.pad(c, ((1, 0), (1, 0)))
        n = 2 * r + 1
        f = (c[n:, n:] - c[:-n, n:] - c[n:, :-n] + c[:-n, :-n]) / (n * n)
    return f / (f.std() + 1e-12)


def make_scene(num_kf=2, W=128, H=96, L=4, F=16, C=8, num_samples=None, mask="full", seed=1234, pose_noise=0.01, with_desc=False,
               step=0.04, rot_step_deg=1.5, back_connections=3):
    """Returns a list of Keyframe (reference layouts, float32).

    num_samples=None -> dense (all valid pixels in raster order); otherwise a seeded random
    subset like Mapper::BuildFrame (core/mapping/mapper.cpp:1222-1239).
    mask: "full" or "ellipse" (endoscope-style inscribed ellipse)."""
    rng = np.random.default_rng(seed)
    cam = np.array([0.8 * W, 0.8 * W, (W - 1) / 2.0, (H - 1) / 2.0, W, H], dtype=F32)
    cams = camera_pyramid(cam, L)
    offs, SP = level_offsets(cams)
    omega = rng.uniform(2.0, 12.0, size=(F, 3)) * rng.choice([-1.0, 1.0], size=(F, 3))
    phi = rng.uniform(0, 2 * np.pi, size=F)
    if mask == "ellipse":
        yy, xx = np.mgrid[0:H, 0:W]
        m = (((xx - cam[2]) / (0.48 * W)) ** 2 + ((yy - cam[3]) / (0.48 * H)) ** 2 <= 1.0).astype(F32)
    else:
        m = np.ones((H, W), dtype=F32)
    masks = mask_pyramid(m, L)
    # sample points come from the mask eroded by 6 px (mapper.cpp:70-71); dense case: all mask pixels
    loc_all, homo_all = valid_locations(m, cam)
    noise_rng = np.random.default_rng(99)

    kfs = []
    u = (np.arange(W) - cam[2]) / cam[0]
    v = (np.arange(H) - cam[3]) / cam[1]
    rays = np.stack(np.broadcast_arrays(u[None, :], v[:, None], np.ones((H, W))), -1).astype(np.float64)
    for k in range(num_kf):
        # ground-truth trajectory: gentle arc, looking down +z
        a = k - (num_kf - 1) / 2.0
        t_true = np.array([step * a, 0.3 * step * np.sin(0.9 * a), 0.2 * step * np.cos(0.7 * a)])
        R_true = _rodrigues(np.deg2rad(rot_step_deg) * np.array([0.3 * np.sin(0.5 * a), a * 0.5, 0.2 * np.cos(0.4 * a)]))
        d = rays @ R_true.T  # world ray directions, camera centre t_true
        # fixed-point ray / height-field intersection: X = o + s d, solve X_z = g(X_x, X_y)
        s = (2.0 - t_true[2]) / d[..., 2]
        for _ in range(20):
            X = t_true + s[..., None] * d
            s = (_surface(X[..., 0], X[..., 1]) - t_true[2]) / d[..., 2]
        X = t_true + s[..., None] * d
        depth = s  # rays have z=1 in the camera frame, so s is the camera-frame depth
        feat = np.sin(np.einsum("fc,hwc->fhw", omega, X) + phi[:, None, None]).astype(F32)
        pyr, grad = gaussian_pyramid_with_grad(feat, masks)

        krng = np.random.default_rng(seed + 1 + k)
        bias = (depth * (1.0 + 0.05 * _smooth_field(krng, H, W, W / 16.0))).astype(F32).reshape(-1)
        basis = np.stack([_smooth_field(krng, H, W, W / 16.0) for _ in range(C)], 0)
        basis = (0.1 * depth.mean() * basis).astype(F32).reshape(C, H * W)  # physical [C,HW]
        jac = basis.T  # [HW,C] view with strides (1,HW), as the reference hands it over

        if num_samples is None or num_samples >= len(loc_all):
            loc, homo = loc_all, homo_all
            tile = os.environ.get("SAGE_SAMPLE_TILE", "")
            if tile:
                tw, th = (int(v) for v in tile.split("x"))
                yy, xx = loc // W, loc % W
                order = np.lexsort((xx % tw, yy % th, xx // tw, yy // th))
                loc, homo = loc[order], homo[order]
        else:
            sel = np.sort(np.random.default_rng(seed + 1000 + k).permutation(len(loc_all))[:num_samples])
            loc, homo = loc_all[sel], homo_all[sel]

        if k == 0:
            R0, t0 = R_true, t_true
        else:
            dlt = noise_rng.normal(0.0, pose_noise, size=6)
            R0 = _rodrigues(dlt[3:]) @ R_true
            t0 = t_true + dlt[:3] * depth.mean() * 0.1
        kfs.append(Keyframe(
            id=k, pose_wk=(R0.astype(F32), t0.astype(F32)), camera_pyramid=cams, level_offsets=offs, video_mask=m,
            feat_map_pyramid=pyr, feat_map_grad_pyramid=grad, dpt_map_bias=bias, dpt_jac_code=jac,
            code=np.zeros(C, dtype=F32), dpt_scale=1.0, sampled_locations_1d=loc, sampled_locations_homo=homo,
            temporal_connections=[j for j in range(max(0, k - back_connections), k)],
            pose_wk_true=(R_true.astype(F32), t_true.astype(F32)),
            feat_desc=feat if with_desc else None))
    return kfs


def ordered_pairs(kfs, mode="temporal"):
    """Ordered keyframe pairs (i -> j): both directions of every temporal link
    (core/mapping/mapper.cpp:339-375), or all ordered pairs for mode='full'."""
    if mode == "full":
        return [(i, j) for i in range(len(kfs)) for j in range(len(kfs)) if i != j]
    pairs = []
    for kf in kfs:
        for j in kf.temporal_connections:
            pairs.append((kf.id, j))
            pairs.append((j, kf.id))
    return pairs


def make_matches(kf0, kf1, M=256, noise_px=0.5, seed=7):
    """Synthetic keypoint matches for the reprojection factor: true correspondences of random
    KF0 pixels (using ground-truth poses and the KF0 bias depth) plus pixel noise.
    Returns (loc1d int32 [M], homo [M,3], match2d [M,2])."""
    rng = np.random.default_rng(seed + 31 * kf0.id + kf1.id)
    cam = kf0.camera_pyramid[0]
    n = len(kf0.sampled_locations_1d)
    sel = rng.permutation(n)[:M]
    loc = kf0.sampled_locations_1d[sel].astype(np.int32)
    homo = kf0.sampled_locations_homo[sel]
    R0, t0 = kf0.pose_wk_true
    R1, t1 = kf1.pose_wk_true
    d = kf0.dpt_map_bias[loc].astype(np.float64)
    Xw = (homo.astype(np.float64) * d[:, None]) @ R0.T.astype(np.float64) + t0
    X1 = (Xw - t1) @ R1.astype(np.float64)
    uv = np.stack([X1[:, 0] / X1[:, 2] * cam[0] + cam[2], X1[:, 1] / X1[:, 2] * cam[1] + cam[3]], 1)
    uv += rng.normal(0, noise_px, size=uv.shape)
    return loc, homo.astype(F32), uv.astype(F32)

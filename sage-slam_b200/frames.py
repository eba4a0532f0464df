"""Host-side data contract of the hot path: cameras, pyramids and the Keyframe record.

Mirrors the reference types the factor kernels consume:
  * PinholeCamera<float>   common/pinhole_camera.h:43-131       -> 6 floats (fx fy u0 v0 w h)
  * CameraPyramid<float>   common/camera_pyramid.h:18-46        -> camera_pyramid()
  * Frame / Keyframe       core/mapping/frame.h:16-125, keyframe.h:19-61 -> Keyframe (same field names)
  * Mapper::GenerateGaussianPyramidWithGrad  core/mapping/mapper.cpp:1385-1426 -> gaussian_pyramid_with_grad()
  * GenerateValidLocations core/mapping/mapping_utils.h:254-287 -> valid_locations()
All arrays are numpy float32 in the REFERENCE layouts; the CUDA side re-lays them out once per
keyframe (csrc/prep.cu).  numpy only: no torch, no oracle.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

F32 = np.float32


def camera_pyramid(cam, levels):
    """[L,6] float32: level i>0 integer-halves width/height and rescales fx fy u0 v0 by the
    fp32 ratio (ResizeViewport, common/pinhole_camera_impl.h:120-132)."""
    cams = [np.asarray(cam, dtype=F32)]
    for i in range(1, levels):
        fx, fy, u0, v0, w, h = cams[-1]
        nw, nh = F32(int(w) // 2), F32(int(h) // 2)
        xr, yr = F32(nw / w), F32(nh / h)
        cams.append(np.array([fx * xr, fy * yr, u0 * xr, v0 * yr, nw, nh], dtype=F32))
    return np.stack(cams)


def level_offsets(cams):
    """Start of each level inside the concatenated pyramid (mapper.cpp:88-97)."""
    sizes = [int(c[4]) * int(c[5]) for c in cams]
    return np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32), int(sum(sizes))


def spatial_grad(x):
    """Central differences with replicate padding (mapping_utils.h:236-252). x [C,H,W] -> (gx, gy)."""
    p = np.pad(x, ((0, 0), (1, 1), (1, 1)), mode="edge")
    gx = F32(0.5) * (p[:, 1:-1, 2:] - p[:, 1:-1, :-2])
    gy = F32(0.5) * (p[:, 2:, 1:-1] - p[:, :-2, 1:-1])
    return gx, gy


def _gauss_down(x):
    """3x3 [1 2 1]^2/16, stride 2, zero padding 1 (mapper.cpp:99-110). x [C,H,W]."""
    C, H, W = x.shape
    ho, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    p = np.pad(x, ((0, 0), (1, 1), (1, 1)))
    k = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=F32) / F32(16)
    out = np.zeros((C, ho, wo), dtype=F32)
    for a in range(3):
        for b in range(3):
            out += k[a, b] * p[:, a:a + 2 * ho:2, b:b + 2 * wo:2]
    return out


def mask_pyramid(mask, levels):
    """Nearest-neighbour halving of the validity mask (mapping_utils.cpp:321-342). mask [H,W]."""
    out = [np.asarray(mask, dtype=F32)]
    for _ in range(levels - 1):
        m = out[-1]
        h, w = m.shape[0] // 2, m.shape[1] // 2
        ys = (np.arange(h) * (m.shape[0] / h)).astype(np.int64)
        xs = (np.arange(w) * (m.shape[1] / w)).astype(np.int64)
        out.append(m[np.ix_(ys, xs)])
    return out


def gaussian_pyramid_with_grad(feat, masks):
    """feat [F,H,W], masks from mask_pyramid() -> (feat_map_pyramid [F,SP], feat_map_grad_pyramid [2,F,SP])."""
    cur = np.asarray(feat, dtype=F32)
    C = cur.shape[0]
    feats, gxs, gys = [cur.reshape(C, -1)], [], []
    gx, gy = spatial_grad(cur)
    gxs.append(gx.reshape(C, -1))
    gys.append(gy.reshape(C, -1))
    for i in range(len(masks) - 1):
        m = masks[i][None]
        cur = _gauss_down(cur * m) / (_gauss_down(m) + F32(1.0e-8))
        gx, gy = spatial_grad(cur)
        feats.append(cur.reshape(C, -1))
        gxs.append(gx.reshape(C, -1))
        gys.append(gy.reshape(C, -1))
    pyr = np.ascontiguousarray(np.concatenate(feats, 1))
    grad = np.ascontiguousarray(np.stack([np.concatenate(gxs, 1), np.concatenate(gys, 1)], 0))
    return pyr, grad


def valid_locations(mask, cam):
    """idx = v*W+u of mask>0.5 pixels and their homogeneous rays ((u-u0)/fx, (v-v0)/fy, 1)."""
    loc1d = np.nonzero(np.asarray(mask).reshape(-1) > 0.5)[0].astype(np.int64)
    W = F32(cam[4])
    x = np.fmod(loc1d.astype(F32), W)
    y = np.floor(loc1d.astype(F32) / W)
    homo = np.stack([(x - F32(cam[2])) / F32(cam[0]), (y - F32(cam[3])) / F32(cam[1]), np.ones_like(x)], 1)
    return loc1d, homo.astype(F32)


@dataclass
class Keyframe:
    """Per-keyframe tensors, field names as in core/mapping/frame.h:16-125.

    pose_wk is (R [3,3], t [3]) keyframe->world.  dpt_jac_code is the [HW, C] *view* of the
    depth net's [C,H,W] output (strides (1, HW)) exactly like the reference
    (core/network/code_depth_network.cpp:38-39)."""
    id: int
    pose_wk: tuple
    camera_pyramid: np.ndarray          # [L,6]
    level_offsets: np.ndarray           # [L] int32
    video_mask: np.ndarray              # [H,W] float 0/1   (*video_mask_ptr)
    feat_map_pyramid: np.ndarray        # [F,SP]
    feat_map_grad_pyramid: np.ndarray   # [2,F,SP]
    dpt_map_bias: np.ndarray            # [HW]
    dpt_jac_code: np.ndarray            # [HW,C] view, strides (1,HW)
    code: np.ndarray                    # [C]
    dpt_scale: float
    sampled_locations_1d: np.ndarray    # [N] int64
    sampled_locations_homo: np.ndarray  # [N,3]
    temporal_connections: List[int] = field(default_factory=list)
    pose_wk_true: Optional[tuple] = None
    feat_desc: Optional[np.ndarray] = None      # [Cd,H,W] descriptor map (Frame::feat_desc, frame.h:86)
    dpt_map_stored: Optional[np.ndarray] = None  # [H,W] as last written by the mapper's UpdateMap (mapper.cpp:1168)

    @property
    def dpt_map(self):
        """Frame::dpt_map: the map UpdateMap stored, else UpdateDepth (mapping_utils.h:216-222) = scale * (bias + jac . code)."""
        if self.dpt_map_stored is not None:
            return self.dpt_map_stored
        H, W = self.video_mask.shape
        return (F32(self.dpt_scale) * (self.dpt_map_bias + self.dpt_jac_code @ self.code)).reshape(H, W).astype(F32)


def std_shuffle(n, seed, libstdcxx="13"):
    """std::shuffle(iota(n), std::mt19937(seed)) as libstdc++ implements it -- the keypoint / sample draw of the reference
    (core/gtsam/reprojection_factor.cpp:43-50, core/system/camera_tracker.cpp:818-823, core/mapping/mapper.cpp:1222-1239).

    libstdcxx="13": bits/stl_algo.h + uniform_int_distribution of GCC 11+ (Lemire's multiply-shift for a 32-bit engine);
    verified against g++ 13.3 (tests/golden/std_shuffle_gcc13.txt, generated by tests/golden/make_std_shuffle.cpp).
    libstdcxx="9": the same shuffle with GCC <= 10's scale-and-reject uniform_int_distribution (the reference's Docker image,
    nvcr.io/nvidia/pytorch:21.04, ships GCC 9); restated from the sources, not verifiable in this image.
    Returns the permuted indices (int64)."""
    raw = np.random.RandomState(int(seed) & 0xFFFFFFFF)._bit_generator  # init_genrand(seed) == std::mt19937(seed)
    pool = []

    def g():
        if not pool:
            pool.extend(int(v) for v in raw.random_raw(4096)[::-1])
        return pool.pop()

    def uniform(hi):  # uniform integer in [0, hi], hi < 2^32
        erange = hi + 1
        if erange > 0xFFFFFFFF:
            return g()
        if libstdcxx == "13":
            prod = g() * erange
            low = prod & 0xFFFFFFFF
            if low < erange:
                thr = ((1 << 32) - erange) % erange
                while low < thr:
                    prod = g() * erange
                    low = prod & 0xFFFFFFFF
            return prod >> 32
        scaling = 0xFFFFFFFF // erange
        past = erange * scaling
        while True:
            r = g()
            if r < past:
                return r // scaling

    idx = list(range(n))
    if n < 2:
        return np.array(idx, np.int64)
    if 0xFFFFFFFF // n >= n:  # two swaps per engine call
        i = 1
        if n % 2 == 0:
            j = uniform(1)
            idx[i], idx[j] = idx[j], idx[i]
            i += 1
        while i < n:
            b0, b1 = i + 1, i + 2
            x = uniform(b0 * b1 - 1)
            for j in (x // b1, x % b1):
                idx[i], idx[j] = idx[j], idx[i]
                i += 1
    else:
        for i in range(1, n):
            j = uniform(i)
            idx[i], idx[j] = idx[j], idx[i]
    return np.array(idx, np.int64)

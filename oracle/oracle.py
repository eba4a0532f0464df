"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.

Two layers:
  * ctypes bindings to oracle/libsage_oracle.so (sage_oracle.c / oracle_body.inc): the
    factor kernels a1-a5 of SURVEY.md section 8, restated from
    /root/reference/system/sources/cuda/*.cpp.
  * numpy/torch restatements of the host-side pieces around them: camera pyramid
    (common/camera_pyramid.h:18-32), Gaussian pyramid + gradients
    (core/mapping/mapper.cpp:1385-1426, mapping_utils.h:236-252), valid locations
    (mapping_utils.h:254-287), SE(3) exp / retract (mapping_utils.h:316-346,
    gtsam/gtsam_traits.h:45-70), NearestPsd (mapping_utils.h:104-128), the tracker's LM
    loop (core/system/camera_tracker.cpp:1156-1279), the descriptor cycle-matching of the
    factor constructors (core/gtsam/reprojection_factor.cpp:57-92; pinned by the torch replay
    of those lines in make_golden_desc.py, CPU and B200), UpdateDepth (mapping_utils.h:216-222)
    and a dense fp64 normal-equation solve that stands in for the (unbuildable) GTSAM solve
    (solver-level parity is therefore UNPINNED by the reference).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(HERE, "libsage_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("sage_oracle.c", "oracle_body.inc", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


# --------------------------------------------------------------------------------------
# ctypes plumbing
# --------------------------------------------------------------------------------------
def _suffix(dtype):
    return "_f32" if np.dtype(dtype) == np.float32 else "_f64"


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _r(a, dtype):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


def _s(v, dtype):
    return ctypes.c_float(float(v)) if np.dtype(dtype) == np.float32 else ctypes.c_double(float(v))


def _jac_strides(jac):
    """jac0 is [HW, C] with arbitrary strides (the reference's is a (1, HW) permuted view)."""
    item = jac.dtype.itemsize
    return ctypes.c_long(jac.strides[0] // item), ctypes.c_long(jac.strides[1] // item)


def _out(D):
    return np.zeros((D, D), np.float64), np.zeros((D,), np.float64), ctypes.c_double(0), ctypes.c_double(0)


def _cams(cams, dtype):
    return _r(np.asarray(cams, dtype=np.float64).reshape(-1, 6), dtype)


def photometric_jac_error(R10, t10, R0, t0, R1, t1, bias0, jac0, code0, mask1, loc1d, homo, feat0, feat1, grad1,
                          level_offsets, scale0, cams, eps, weights, dtype=np.float32):
    """df::photometric_jac_error_calculate<CS,FS>; returns (AtA, Atb, error, n_inliers)."""
    f = getattr(lib(), "oracle_photometric_jac_error" + _suffix(dtype))
    jac0 = np.asarray(jac0, dtype=dtype)
    N, CS = len(loc1d), jac0.shape[1]
    feat0, feat1, grad1 = _r(feat0, dtype), _r(feat1, dtype), _r(grad1, dtype)
    FS, SP = feat0.shape
    cams = _cams(cams, dtype)
    L = cams.shape[0]
    A, b, e, n = _out(13 + CS)
    sr, sc = _jac_strides(jac0)
    args = [_r(x, dtype) for x in (R10, t10, R0, t0, R1, t1, bias0)]
    code0, mask1, homo, weights = _r(code0, dtype), _r(mask1, dtype), _r(homo, dtype), _r(weights, dtype)
    loc1d = _r(loc1d, np.int64)
    lo = _r(level_offsets, np.int32)
    f(_p(A), _p(b), ctypes.byref(e), ctypes.byref(n), *[_p(x) for x in args], _p(jac0), sr, sc, _p(code0), _p(mask1),
      _p(loc1d), _p(homo), _p(feat0), _p(feat1), _p(grad1), _p(lo), _s(scale0, dtype), _p(cams), _s(eps, dtype),
      _p(weights), ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(FS), ctypes.c_int(CS), ctypes.c_long(SP))
    return A, b, e.value, n.value


def photometric_error(R10, t10, bias0, jac0, code0, mask1, loc1d, homo, feat0, feat1, level_offsets, scale0, cams,
                      eps, weights, dtype=np.float32):
    f = getattr(lib(), "oracle_photometric_error" + _suffix(dtype))
    jac0 = np.asarray(jac0, dtype=dtype)
    N, CS = len(loc1d), jac0.shape[1]
    feat0, feat1 = _r(feat0, dtype), _r(feat1, dtype)
    FS, SP = feat0.shape
    cams = _cams(cams, dtype)
    L = cams.shape[0]
    e, n = ctypes.c_double(0), ctypes.c_double(0)
    sr, sc = _jac_strides(jac0)
    R10, t10, bias0, code0, mask1, homo, weights = [_r(x, dtype) for x in (R10, t10, bias0, code0, mask1, homo, weights)]
    loc1d = _r(loc1d, np.int64)
    lo = _r(level_offsets, np.int32)
    f(ctypes.byref(e), ctypes.byref(n), _p(R10), _p(t10), _p(bias0), _p(jac0), sr, sc, _p(code0), _p(mask1), _p(loc1d),
      _p(homo), _p(feat0), _p(feat1), _p(lo), _s(scale0, dtype), _p(cams), _s(eps, dtype), _p(weights),
      ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(FS), ctypes.c_int(CS), ctypes.c_long(SP))
    return e.value, n.value


def tracker_photo_jac_error(R, t, mask1, dpts0, homo, sfeat0, feat1, grad1, level_offsets, cams, eps, weights,
                            scale0=None, dtype=np.float32):
    """tracker_photo_jac_error_calculate (scale0 None) or ..._with_scale (scale0 given)."""
    f = getattr(lib(), "oracle_tracker_photo_jac_error" + _suffix(dtype))
    with_scale = scale0 is not None
    D = 7 if with_scale else 6
    feat1, grad1, sfeat0 = _r(feat1, dtype), _r(grad1, dtype), _r(sfeat0, dtype)
    FS, SP = feat1.shape
    cams = _cams(cams, dtype)
    L, N = cams.shape[0], len(dpts0)
    A, b, e, n = _out(D)
    R, t, mask1, dpts0, homo, weights = [_r(x, dtype) for x in (R, t, mask1, dpts0, homo, weights)]
    lo = _r(level_offsets, np.int32)
    f(_p(A), _p(b), ctypes.byref(e), ctypes.byref(n), _p(R), _p(t), _p(mask1), _p(dpts0), _p(homo), _p(sfeat0),
      _p(feat1), _p(grad1), _p(lo), _p(cams), _s(scale0 if with_scale else 1.0, dtype), _s(eps, dtype), _p(weights),
      ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(FS), ctypes.c_long(SP), ctypes.c_int(int(with_scale)))
    return A, b, e.value, n.value


def tracker_photo_error(R, t, mask1, dpts0, homo, sfeat0, feat1, level_offsets, cams, eps, weights, dtype=np.float32):
    f = getattr(lib(), "oracle_tracker_photo_error" + _suffix(dtype))
    feat1, sfeat0 = _r(feat1, dtype), _r(sfeat0, dtype)
    FS, SP = feat1.shape
    cams = _cams(cams, dtype)
    L, N = cams.shape[0], len(dpts0)
    e, n = ctypes.c_double(0), ctypes.c_double(0)
    R, t, mask1, dpts0, homo, weights = [_r(x, dtype) for x in (R, t, mask1, dpts0, homo, weights)]
    lo = _r(level_offsets, np.int32)
    f(ctypes.byref(e), ctypes.byref(n), _p(R), _p(t), _p(mask1), _p(dpts0), _p(homo), _p(sfeat0), _p(feat1), _p(lo),
      _p(cams), _s(eps, dtype), _p(weights), ctypes.c_int(N), ctypes.c_int(L), ctypes.c_int(FS), ctypes.c_long(SP))
    return e.value, n.value


def geometric_jac_error(R10, t10, R0, t0, R1, t1, bias0, jac0, code0, dpt1, dgrad1, basis1, mask1, loc1d, homo,
                        scale0, scale1, cam, eps, loss_param, weight, dtype=np.float32):
    f = getattr(lib(), "oracle_geometric_jac_error" + _suffix(dtype))
    jac0 = np.asarray(jac0, dtype=dtype)
    N, CS = len(loc1d), jac0.shape[1]
    A, b, e, n = _out(14 + 2 * CS)
    sr, sc = _jac_strides(jac0)
    a = [_r(x, dtype) for x in (R10, t10, R0, t0, R1, t1, bias0)]
    code0, dpt1, dgrad1, basis1, mask1, homo = [_r(x, dtype) for x in (code0, dpt1, dgrad1, basis1, mask1, homo)]
    loc1d = _r(loc1d, np.int32)
    cam = _cams(cam, dtype)
    f(_p(A), _p(b), ctypes.byref(e), ctypes.byref(n), *[_p(x) for x in a], _p(jac0), sr, sc, _p(code0), _p(dpt1),
      _p(dgrad1), _p(basis1), _p(mask1), _p(loc1d), _p(homo), _s(scale0, dtype), _s(scale1, dtype), _p(cam),
      _s(eps, dtype), _s(loss_param, dtype), _s(weight, dtype), ctypes.c_int(N), ctypes.c_int(CS))
    return A, b, e.value, n.value


def geometric_error(R10, t10, bias0, jac0, code0, dpt1, mask1, loc1d, homo, scale0, cam, eps, loss_param, weight,
                    dtype=np.float32):
    f = getattr(lib(), "oracle_geometric_error" + _suffix(dtype))
    jac0 = np.asarray(jac0, dtype=dtype)
    N, CS = len(loc1d), jac0.shape[1]
    e, n = ctypes.c_double(0), ctypes.c_double(0)
    sr, sc = _jac_strides(jac0)
    R10, t10, bias0, code0, dpt1, mask1, homo = [_r(x, dtype) for x in (R10, t10, bias0, code0, dpt1, mask1, homo)]
    loc1d = _r(loc1d, np.int32)
    cam = _cams(cam, dtype)
    f(ctypes.byref(e), ctypes.byref(n), _p(R10), _p(t10), _p(bias0), _p(jac0), sr, sc, _p(code0), _p(dpt1), _p(mask1),
      _p(loc1d), _p(homo), _s(scale0, dtype), _p(cam), _s(eps, dtype), _s(loss_param, dtype), _s(weight, dtype),
      ctypes.c_int(N), ctypes.c_int(CS))
    return e.value, n.value


def reprojection_jac_error(R10, t10, R0, t0, R1, t1, bias0, jac0, code0, loc1d, homo, match2d, scale0, cam, eps,
                           loss_param, weight, dtype=np.float32):
    f = getattr(lib(), "oracle_reprojection_jac_error" + _suffix(dtype))
    jac0 = np.asarray(jac0, dtype=dtype)
    M, CS = len(loc1d), jac0.shape[1]
    A, b, e, n = _out(13 + CS)
    sr, sc = _jac_strides(jac0)
    a = [_r(x, dtype) for x in (R10, t10, R0, t0, R1, t1, bias0)]
    code0, homo, match2d = [_r(x, dtype) for x in (code0, homo, match2d)]
    loc1d = _r(loc1d, np.int32)
    cam = _cams(cam, dtype)
    f(_p(A), _p(b), ctypes.byref(e), ctypes.byref(n), *[_p(x) for x in a], _p(jac0), sr, sc, _p(code0), _p(loc1d),
      _p(homo), _p(match2d), _s(scale0, dtype), _p(cam), _s(eps, dtype), _s(loss_param, dtype), _s(weight, dtype),
      ctypes.c_int(M), ctypes.c_int(CS))
    return A, b, e.value, n.value


def reprojection_error(R10, t10, bias0, jac0, code0, loc1d, homo, match2d, scale0, cam, eps, loss_param, weight,
                       dtype=np.float32):
    f = getattr(lib(), "oracle_reprojection_error" + _suffix(dtype))
    jac0 = np.asarray(jac0, dtype=dtype)
    M, CS = len(loc1d), jac0.shape[1]
    e, n = ctypes.c_double(0), ctypes.c_double(0)
    sr, sc = _jac_strides(jac0)
    R10, t10, bias0, code0, homo, match2d = [_r(x, dtype) for x in (R10, t10, bias0, code0, homo, match2d)]
    loc1d = _r(loc1d, np.int32)
    cam = _cams(cam, dtype)
    f(ctypes.byref(e), ctypes.byref(n), _p(R10), _p(t10), _p(bias0), _p(jac0), sr, sc, _p(code0), _p(loc1d), _p(homo),
      _p(match2d), _s(scale0, dtype), _p(cam), _s(eps, dtype), _s(loss_param, dtype), _s(weight, dtype),
      ctypes.c_int(M), ctypes.c_int(CS))
    return e.value, n.value


def tracker_reproj_jac_error(R, t, dpts0, homo, match2d, cam, eps, loss_param, weight, dtype=np.float32):
    f = getattr(lib(), "oracle_tracker_reproj_jac_error" + _suffix(dtype))
    M = len(dpts0)
    A, b, e, n = _out(6)
    R, t, dpts0, homo, match2d = [_r(x, dtype) for x in (R, t, dpts0, homo, match2d)]
    cam = _cams(cam, dtype)
    f(_p(A), _p(b), ctypes.byref(e), ctypes.byref(n), _p(R), _p(t), _p(dpts0), _p(homo), _p(match2d), _p(cam),
      _s(eps, dtype), _s(loss_param, dtype), _s(weight, dtype), ctypes.c_int(M))
    return A, b, e.value, n.value


def tracker_reproj_error(R, t, dpts0, homo, match2d, cam, eps, loss_param, weight, dtype=np.float32):
    f = getattr(lib(), "oracle_tracker_reproj_error" + _suffix(dtype))
    M = len(dpts0)
    e, n = ctypes.c_double(0), ctypes.c_double(0)
    R, t, dpts0, homo, match2d = [_r(x, dtype) for x in (R, t, dpts0, homo, match2d)]
    cam = _cams(cam, dtype)
    f(ctypes.byref(e), ctypes.byref(n), _p(R), _p(t), _p(dpts0), _p(homo), _p(match2d), _p(cam), _s(eps, dtype),
      _s(loss_param, dtype), _s(weight, dtype), ctypes.c_int(M))
    return e.value, n.value



def tracker_match_geom_jac_error(R, t, dpts0, dpts1, homo0, homo1, loss_param, weight, scale0=None, dtype=np.float32):
    """tracker_match_geom_jac_error_calculate (scale0 None, 6x6) or ..._with_scale (7x7); dpts0 is the SCALED depth."""
    f = getattr(lib(), "oracle_tracker_match_geom_jac_error" + _suffix(dtype))
    with_scale = scale0 is not None
    D = 7 if with_scale else 6
    A, b, e, _ = _out(D)
    R, t, dpts0, dpts1, homo0, homo1 = [_r(x, dtype) for x in (R, t, dpts0, dpts1, homo0, homo1)]
    f(_p(A), _p(b), ctypes.byref(e), _p(R), _p(t), _p(dpts0), _p(dpts1), _p(homo0), _p(homo1),
      _s(scale0 if with_scale else 1.0, dtype), _s(loss_param, dtype), _s(weight, dtype), ctypes.c_int(len(dpts0)),
      ctypes.c_int(int(with_scale)))
    return A, b, e.value


def tracker_match_geom_error(R, t, dpts0, dpts1, homo0, homo1, loss_param, weight, dtype=np.float32):
    f = getattr(lib(), "oracle_tracker_match_geom_error" + _suffix(dtype))
    e = ctypes.c_double(0)
    R, t, dpts0, dpts1, homo0, homo1 = [_r(x, dtype) for x in (R, t, dpts0, dpts1, homo0, homo1)]
    f(ctypes.byref(e), _p(R), _p(t), _p(dpts0), _p(dpts1), _p(homo0), _p(homo1), _s(loss_param, dtype), _s(weight, dtype),
      ctypes.c_int(len(dpts0)))
    return e.value


LOSS_TYPES = {"fair": 0, "L2": 1, "huber": 2, "unbiased": 3}


def _mg_call(fname, with_pose, R10, t10, R0, t0, R1, t1, bias0, bias1, jac0, jac1, code0, code1, loc0, loc1, dpts0, dpts1,
             homo0, homo1, scale0, scale1, loss_param, weight, loss_type, dtype):
    f = getattr(lib(), fname + _suffix(dtype))
    loop = bias0 is None
    M = len(homo0)
    null = ctypes.c_void_p(0)
    zero = ctypes.c_long(0)
    if loop:
        CS = 0
        maps = [null, null, null, zero, zero, null, zero, zero, null, null, null, null]
        dp = [_r(dpts0, dtype), _r(dpts1, dtype)]
        keep = dp
        depth_args = [_p(dp[0]), _p(dp[1])]
    else:
        jac0, jac1 = np.asarray(jac0, dtype=dtype), np.asarray(jac1, dtype=dtype)
        CS = jac0.shape[1]
        b0, b1, c0, c1 = [_r(x, dtype) for x in (bias0, bias1, code0, code1)]
        l0, l1 = _r(loc0, np.int32), _r(loc1, np.int32)
        keep = [b0, b1, c0, c1, l0, l1, jac0, jac1]
        maps = [_p(b0), _p(b1), _p(jac0), *_jac_strides(jac0), _p(jac1), *_jac_strides(jac1), _p(c0), _p(c1), _p(l0), _p(l1)]
        depth_args = [null, null]
    h0, h1 = _r(homo0, dtype), _r(homo1, dtype)
    tail = [*maps, *depth_args, _p(h0), _p(h1), _s(scale0, dtype), _s(scale1, dtype), _s(loss_param, dtype), _s(weight, dtype),
            ctypes.c_int(M), ctypes.c_int(CS), ctypes.c_int(LOSS_TYPES[loss_type] if isinstance(loss_type, str) else int(loss_type))]
    e = ctypes.c_double(0)
    if with_pose:
        D = 14 + 2 * CS
        A, b, _, _ = _out(D)
        poses = [_r(x, dtype) for x in (R10, t10, R0, t0, R1, t1)]
        f(_p(A), _p(b), ctypes.byref(e), *[_p(x) for x in poses], *tail)
        return A, b, e.value
    poses = [_r(x, dtype) for x in (R10, t10)]
    f(ctypes.byref(e), *[_p(x) for x in poses], *tail)
    return e.value


def match_geometry_jac_error(R10, t10, R0, t0, R1, t1, bias0, bias1, jac0, jac1, code0, code1, homo0, homo1, loc0, loc1,
                             scale0, scale1, loss_param, weight, loss_type="fair", dtype=np.float32):
    """df::match_geometry_jac_error_calculate<CS>; order [pose0 pose1 code0 code1 scale0 scale1]."""
    return _mg_call("oracle_match_geometry_jac_error", True, R10, t10, R0, t0, R1, t1, bias0, bias1, jac0, jac1, code0, code1,
                    loc0, loc1, None, None, homo0, homo1, scale0, scale1, loss_param, weight, loss_type, dtype)


def match_geometry_error(R10, t10, bias0, bias1, jac0, jac1, code0, code1, homo0, homo1, loc0, loc1, scale0, scale1,
                         loss_param, weight, loss_type="fair", dtype=np.float32):
    return _mg_call("oracle_match_geometry_error", False, R10, t10, None, None, None, None, bias0, bias1, jac0, jac1, code0,
                    code1, loc0, loc1, None, None, homo0, homo1, scale0, scale1, loss_param, weight, loss_type, dtype)


def loop_mg_jac_error(R10, t10, R0, t0, R1, t1, dpts0, dpts1, homo0, homo1, scale0, scale1, loss_param, weight, dtype=np.float32):
    """df::loop_mg_jac_error_calculate; dpts are UNSCALED depths; order [pose0 pose1 scale0 scale1]."""
    return _mg_call("oracle_match_geometry_jac_error", True, R10, t10, R0, t0, R1, t1, None, None, None, None, None, None,
                    None, None, dpts0, dpts1, homo0, homo1, scale0, scale1, loss_param, weight, 0, dtype)


def loop_mg_error(R10, t10, dpts0, dpts1, homo0, homo1, scale0, scale1, loss_param, weight, dtype=np.float32):
    return _mg_call("oracle_match_geometry_error", False, R10, t10, None, None, None, None, None, None, None, None, None, None,
                    None, None, dpts0, dpts1, homo0, homo1, scale0, scale1, loss_param, weight, 0, dtype)


# --------------------------------------------------------------------------------------
# host-side restatements
# --------------------------------------------------------------------------------------
def cycle_match(feat_desc_0, feat_desc_1, keypoint_locations_1d, cyc_consis_thresh, chunk=64):
    """Dense descriptor cycle-matching, restating core/gtsam/reprojection_factor.cpp:57-92 (identical in
    match_geometry_factor.cpp:62-97 and camera_tracker.cpp:798-834) expression by expression:
      response[k, p] = -sum_c square(q[c, k] - desc[c, p])  (fp32; sub, square and the channel sum are separate roundings,
      channels added in order), argmax over p (first maximum), return pass with the matched descriptors, then the
      cycle distance test on fmod(loc, W) / floor(loc / W).
    Keypoints are processed `chunk` at a time only to bound memory; per-element arithmetic is unchanged."""
    d0 = np.asarray(feat_desc_0, np.float32)
    d0 = d0.reshape((-1,) + d0.shape[-2:])
    d1 = np.asarray(feat_desc_1, np.float32).reshape(d0.shape)
    C, H, W = d0.shape
    d0 = d0.reshape(C, H * W)
    d1 = d1.reshape(C, H * W)
    kp = np.asarray(keypoint_locations_1d, np.int64)

    def best(q, desc):  # q [C, K]
        out = np.zeros(q.shape[1], np.int64)
        for k0 in range(0, q.shape[1], chunk):
            qq = q[:, k0:k0 + chunk]
            acc = np.zeros((qq.shape[1], desc.shape[1]), np.float32)
            for c in range(C):
                d = qq[c][:, None] - desc[c][None, :]
                acc += d * d
            out[k0:k0 + chunk] = np.argmax(-acc, axis=1)
        return out

    raw1 = best(d0[:, kp], d1)
    cyc0 = best(d1[:, raw1], d0)
    fw = np.float32(W)
    kx, ky = np.fmod(kp.astype(np.float32), fw), np.floor(kp.astype(np.float32) / fw)
    cx, cy = np.fmod(cyc0.astype(np.float32), fw), np.floor(cyc0.astype(np.float32) / fw)
    dist = np.square(kx - cx) + np.square(ky - cy)
    sel = np.nonzero(dist <= np.float32(cyc_consis_thresh) * np.float32(cyc_consis_thresh))[0]
    loc1 = raw1[sel]
    return {"raw_matched_locations_1d_1": raw1.astype(np.int32), "cyc_matched_locations_1d_0": cyc0.astype(np.int32),
            "inlier_within_keypoint_indexes": sel.astype(np.int64), "matched_locations_1d_1": loc1.astype(np.int32),
            "matched_locations_2d_1": np.stack([np.fmod(loc1.astype(np.float32), fw), np.floor(loc1.astype(np.float32) / fw)], 1)}


def update_depth(dpt_map_bias, dpt_jac_code, code, dpt_scale):
    """UpdateDepth (core/mapping/mapping_utils.h:216-222): dpt_map = scale * (bias + jac . code), fp32."""
    b = np.asarray(dpt_map_bias, np.float32).reshape(-1)
    return (np.float32(dpt_scale) * (b + np.asarray(dpt_jac_code, np.float32) @ np.asarray(code, np.float32))).astype(np.float32)


def camera_pyramid(cam, levels):
    """CameraPyramid<float>(cam, levels): common/camera_pyramid.h:18-32 +
    PinholeCamera::ResizeViewport (common/pinhole_camera_impl.h:120-132).  fp32 arithmetic,
    integer-halved width/height stored back as float."""
    f = np.float32
    cams = [np.array(cam, dtype=f)]
    for i in range(1, levels):
        fx, fy, u0, v0, w, h = cams[i - 1]
        nw, nh = f(int(w) // 2), f(int(h) // 2)  # size_t division of the float width/height
        xr, yr = f(nw / w), f(nh / h)
        cams.append(np.array([fx * xr, fy * yr, u0 * xr, v0 * yr, nw, nh], dtype=f))
    return np.stack(cams)


def level_offsets(cams):
    """core/mapping/mapper.cpp:88-97"""
    off, out = 0, []
    for c in cams:
        out.append(off)
        off += int(c[4]) * int(c[5])
    return np.array(out, np.int32), off


def spatial_grad(x):
    """ComputeSpatialGrad (core/mapping/mapping_utils.h:236-252): central differences with
    replicate padding.  x: torch [1,C,H,W] -> [1,2C,H,W] (all d/dx first, then all d/dy)."""
    import torch
    import torch.nn.functional as F

    H, W = x.shape[2], x.shape[3]
    p = F.pad(x, (1, 1, 1, 1), mode="replicate")
    gx = 0.5 * (p[:, :, 1:H + 1, 2:W + 2] - p[:, :, 1:H + 1, 0:W])
    gy = 0.5 * (p[:, :, 2:H + 2, 1:W + 1] - p[:, :, 0:H, 1:W + 1])
    return torch.cat([gx, gy], 1)


def mask_pyramid(mask, levels):
    """GenerateMaskPyramid (core/mapping/mapping_utils.cpp:321-342): nearest resize by halving."""
    import torch.nn.functional as F

    out = [mask]
    h, w = mask.shape[2], mask.shape[3]
    cur = mask
    for _ in range(levels - 1):
        h //= 2
        w //= 2
        cur = F.interpolate(cur, size=(h, w), mode="nearest")
        out.append(cur)
    return out


def gaussian_pyramid_with_grad(feat, masks):
    """Mapper::GenerateGaussianPyramidWithGrad (core/mapping/mapper.cpp:1385-1426).
    feat: torch [1,F,H,W]; masks: list of [1,1,h,w].  Kernel: 3x3 [1 2 1]^2/16, stride 2, pad 1
    (mapper.cpp:99-110).  Returns feat_pyr [F, SP], grad_pyr [2, F, SP]."""
    import torch
    import torch.nn.functional as F

    C, H, W = feat.shape[1], feat.shape[2], feat.shape[3]
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], dtype=feat.dtype).reshape(1, 1, 3, 3) / 16.0
    cur = feat.reshape(C, 1, H, W)
    feats = [cur.reshape(C, H * W)]
    grads = [spatial_grad(cur.reshape(1, C, H, W)).reshape(2, C, H * W)]
    for i in range(len(masks) - 1):
        m = masks[i]
        raw = F.conv2d(cur * m, k, stride=2, padding=1)
        rm = F.conv2d(m, k, stride=2, padding=1)
        h, w = raw.shape[2], raw.shape[3]
        cur = raw / (rm + 1.0e-8)
        feats.append(cur.reshape(C, h * w))
        grads.append(spatial_grad(cur.reshape(1, C, h, w)).reshape(2, C, h * w))
    return torch.cat(feats, 1).contiguous(), torch.cat(grads, 2).contiguous()


def valid_locations(mask, cam):
    """GenerateValidLocations (core/mapping/mapping_utils.h:254-287): idx = v*W+u,
    homo = ((u-u0)/fx, (v-v0)/fy, 1) in fp32."""
    f = np.float32
    m = np.asarray(mask).reshape(-1)
    loc1d = np.nonzero(m > 0.5)[0].astype(np.int64)
    W = f(cam[4])
    x = np.fmod(loc1d.astype(f), W)
    y = np.floor(loc1d.astype(f) / W)
    homo = np.stack([(x - f(cam[2])) / f(cam[0]), (y - f(cam[3])) / f(cam[1]), np.ones_like(x)], 1).astype(f)
    return loc1d, homo


def so3_hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=w.dtype)


def se3_exp(omega, v):
    """se3_exp (core/mapping/mapping_utils.h:316-346); dtype follows the inputs."""
    omega = np.asarray(omega)
    v = np.asarray(v, dtype=omega.dtype)
    dt = omega.dtype.type
    theta = dt(np.linalg.norm(omega))
    n = omega / theta if theta > 0 else np.array([1, 0, 0], dtype=omega.dtype)
    theta = max(theta, dt(1.0e-14))
    s, c = dt(np.sin(theta)), dt(np.cos(theta))
    K = so3_hat(n)
    K2 = K @ K
    I = np.eye(3, dtype=omega.dtype)
    R = I + s * K + (dt(1) - c) * K2
    V = I + ((dt(1) - c) / theta) * K + ((theta - s) / theta) * K2
    return R, V @ v


def retract(R, t, delta):
    """Left-multiplicative SE(3) retraction, delta = [v, omega]
    (core/gtsam/gtsam_traits.h:45-70; camera_tracker.cpp:491-512)."""
    dR, dt_ = se3_exp(np.asarray(delta[3:6]), np.asarray(delta[0:3]))
    return dR @ R, dR @ t + dt_


def relative_pose(R0, t0, R1, t1):
    """T10 = T1^-1 T0 as the factors compute it (core/gtsam/photometric_factor.cpp:280-281)."""
    return R1.T @ R0, R1.T @ (t0 - t1)


def rotation_to_angle_axis(R, eps=1.0e-6):
    """RotationToAngleAxis (core/mapping/mapping_utils.h:145-214) for the generic branch
    structure of the reference (quaternion via the transposed matrix, then angle-axis)."""
    m = np.asarray(R, dtype=np.float64).T
    d2 = m[2, 2] < eps
    d0_d1 = m[0, 0] > m[1, 1]
    d0_nd1 = m[0, 0] < -m[1, 1]
    t0 = 1 + m[0, 0] - m[1, 1] - m[2, 2]
    q0 = np.array([m[1, 2] - m[2, 1], t0, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]])
    t1 = 1 - m[0, 0] + m[1, 1] - m[2, 2]
    q1 = np.array([m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t1, m[1, 2] + m[2, 1]])
    t2 = 1 - m[0, 0] - m[1, 1] + m[2, 2]
    q2 = np.array([m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t2])
    t3 = 1 + m[0, 0] + m[1, 1] + m[2, 2]
    q3 = np.array([t3, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]])
    c0, c1 = d2 and d0_d1, d2 and not d0_d1
    c2, c3 = (not d2) and d0_nd1, (not d2) and not d0_nd1
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    # NOTE reference quirk (mapping_utils.h:187): t0 is multiplied by mask_c1, not mask_c0.
    q = 0.5 * q / np.sqrt(t0 * c1 + t1 * c1 + t2 * c2 + t3 * c3)
    s2 = q[1] ** 2 + q[2] ** 2 + q[3] ** 2
    s = np.sqrt(s2)
    two_theta = np.arctan2(-s, -q[0]) if q[0] < 0 else np.arctan2(s, q[0])
    k = two_theta / s if s2 > 0 else 2.0
    return k * q[1:4]


def _make_jacobi(x, y, z):
    """JacobiRotation::makeJacobi(x, y, z) (Eigen 3.3.9, Jacobi/Jacobi.h:83-113): (c, s)."""
    tiny = np.finfo(np.float64).tiny
    deno = 2.0 * abs(y)
    if deno < tiny:
        return 1.0, 0.0
    tau = (x - z) / deno
    w = np.sqrt(tau * tau + 1.0)
    t = 1.0 / (tau + w) if tau > 0 else 1.0 / (tau - w)
    sign_t = 1.0 if t > 0 else -1.0
    n = 1.0 / np.sqrt(t * t + 1.0)
    return n, -sign_t * (y / abs(y)) * abs(t) * n


def eigen_jacobi_svd_v(B):
    """Singular values (descending) and V of a square real matrix exactly as Eigen 3.3.9's JacobiSVD<MatrixXd>(B, ComputeThinV)
    produces them (SVD/JacobiSVD.h `compute`, misc/RealSvd2x2.h): two-sided Jacobi sweeps over (p, q), q < p, rotations
    accumulated into V, selection sort by singular value.  The SIGNS of V's columns are the algorithm's, not a convention --
    and the reference's NearestPsd depends on them (V^T S V is not invariant under column sign flips), which is why numpy's
    LAPACK SVD cannot stand in for it."""
    B = np.asarray(B, np.float64)
    n = B.shape[0]
    tiny = np.finfo(np.float64).tiny
    precision = 2.0 * np.finfo(np.float64).eps
    scale = np.abs(B).max()
    if scale == 0:
        scale = 1.0
    W = B / scale
    V = np.eye(n)
    max_diag = np.abs(np.diag(W)).max()
    finished = False
    while not finished:
        finished = True
        for p in range(1, n):
            for q in range(p):
                thr = max(tiny, precision * max_diag)
                if abs(W[p, q]) > thr or abs(W[q, p]) > thr:
                    finished = False
                    m00, m01, m10, m11 = W[p, p], W[p, q], W[q, p], W[q, q]
                    t, d = m00 + m11, m10 - m01
                    if abs(d) < tiny:
                        rs, rc = 0.0, 1.0
                    else:
                        u = t / d
                        tmp = np.sqrt(1.0 + u * u)
                        rs, rc = 1.0 / tmp, u / tmp
                    # m.applyOnTheLeft(0, 1, rot1): x' = c x + s y, y' = -s x + c y on the two rows
                    a00, a01 = rc * m00 + rs * m10, rc * m01 + rs * m11
                    a11 = -rs * m01 + rc * m11
                    cr, sr = _make_jacobi(a00, a01, a11)          # j_right
                    cl, sl = rc * cr + rs * sr, rs * cr - rc * sr  # j_left = rot1 * j_right.transpose()
                    xp, xq = W[p, :].copy(), W[q, :].copy()        # applyOnTheLeft(p, q, j_left)
                    W[p, :], W[q, :] = cl * xp + sl * xq, -sl * xp + cl * xq
                    xp, xq = W[:, p].copy(), W[:, q].copy()        # applyOnTheRight(p, q, j_right): rotation (c, -s) on the columns
                    W[:, p], W[:, q] = cr * xp - sr * xq, sr * xp + cr * xq
                    xp, xq = V[:, p].copy(), V[:, q].copy()
                    V[:, p], V[:, q] = cr * xp - sr * xq, sr * xp + cr * xq
                    max_diag = max(max_diag, abs(W[p, p]), abs(W[q, q]))
    sv = np.abs(np.diag(W)) * scale
    for i in range(n):
        pos = int(np.argmax(sv[i:]))
        if sv[i + pos] == 0:
            break
        if pos:
            pos += i
            sv[[i, pos]] = sv[[pos, i]]
            V[:, [i, pos]] = V[:, [pos, i]]
    return sv, V


def eigen_ldlt_is_positive(M):
    """Eigen::LDLT<MatrixXd>(M).isPositive() (Eigen 3.3.9 Cholesky/LDLT.h `unblocked`): pivoted LDL^T of the lower triangle,
    the sign is read off the pivots as they appear (no tolerance), so a rank-deficient PSD matrix can report `false`."""
    mat = np.array(M, np.float64)
    n = mat.shape[0]
    if n == 0:
        return True
    if n == 1:
        return mat[0, 0] >= 0
    sign = 0  # 0 zero, +1 positive semi-definite, -1 negative semi-definite, 2 indefinite
    for k in range(n):
        b = k + int(np.argmax(np.abs(np.diag(mat)[k:])))
        if b != k:
            s = n - b - 1
            mat[[k, b], :k] = mat[[b, k], :k]
            if s:
                mat[b + 1:, [k, b]] = mat[b + 1:, [b, k]]
            mat[k, k], mat[b, b] = mat[b, b], mat[k, k]
            for i in range(k + 1, b):
                mat[i, k], mat[b, i] = mat[b, i], mat[i, k]
        rs = n - k - 1
        if k > 0:
            temp = np.diag(mat)[:k] * mat[k, :k]
            mat[k, k] -= mat[k, :k] @ temp
            if rs > 0:
                mat[k + 1:, k] -= mat[k + 1:, :k] @ temp
        akk = mat[k, k]
        valid = abs(akk) > 0
        if k == 0 and not valid:
            return True  # ZeroSign
        if rs > 0 and valid:
            mat[k + 1:, k] /= akk
        if sign == 1:
            if akk < 0:
                sign = 2
        elif sign == -1:
            if akk > 0:
                sign = 2
        elif sign == 0:
            sign = 1 if akk > 0 else (-1 if akk < 0 else 0)
    return sign in (1, 0)


def eigen_nearest_psd(M):
    """df::NearestPsd (core/mapping/mapping_utils.h:104-128) with Eigen 3.3.9's JacobiSVD / LDLT behaviour restated, so that the
    NUMBERS (not only the formula) are the reference's: pinned by tests/golden/host_pins.npz, which the reference's own code
    compiled against its vendored Eigen produced (oracle/make_golden_host.py)."""
    M = np.asarray(M, np.float64)
    B = (M + M.T) / 2
    sv, V = eigen_jacobi_svd_v(B)
    H = V.T @ np.diag(sv) @ V  # sic: V^T S V
    A2 = (B + H) / 2
    A3 = (A2 + A2.T) / 2
    k, I = 1, np.eye(M.shape[0])
    while not eigen_ldlt_is_positive(A3):
        A3 = A3 + I * (-np.linalg.eigvalsh(A3).min() * k + 1e-15)
        k *= 2
    return A3


def nearest_psd(M, reference_faithful=True):
    """NearestPsd (core/mapping/mapping_utils.h:104-128).  reference_faithful=True reproduces the reference's NUMBERS: its
    V^T S V with the column signs Eigen 3.3.9's JacobiSVD yields and Eigen's LDLT positivity test
    (eigen_nearest_psd; pinned by tests/golden/host_pins.npz = the reference's own code run against its vendored Eigen).
    False is Higham's projection V S V^T."""
    M = np.asarray(M, dtype=np.float64)
    if reference_faithful:
        return eigen_nearest_psd(M)
    B = (M + M.T) / 2
    _, s, Vt = np.linalg.svd(B)
    V = Vt.T
    A2 = (B + V @ np.diag(s) @ V.T) / 2
    A3 = (A2 + A2.T) / 2
    k = 1
    I = np.eye(M.shape[0])

    def is_psd(A):
        try:
            np.linalg.cholesky(A)
            return True
        except np.linalg.LinAlgError:
            return False

    while not is_psd(A3):
        A3 = A3 + I * (-np.linalg.eigvalsh(A3).min() * k + 1e-15)
        k *= 2
    return A3


def tracker_lm7(jac_fn, err_fn, R, t, scale, init_damp=1e-4, min_damp=1e-6, max_damp=1e-2, damp_dec=10.0, damp_inc=100.0,
                max_iters=40, jac_thresh=1e-2, min_grad=1e-8, min_param_inc=1e-8):
    """CameraTracker::TrackFrame 7-DoF LM loop (core/system/camera_tracker.cpp:1479-1630): relative pose + scale_0.
    jac_fn(R,t,s)->(AtA 7x7, Atb 7, err), err_fn(R,t,s)->err."""
    f = np.float32
    # the reference's options are floats (camera_tracker.h:54-63): compare float with float, or `damp < max_damp` never turns false
    # for a max_damp that fp32 rounds down (1e-2) and the loop cannot stop at maximum damping
    init_damp, min_damp, max_damp, damp_dec, damp_inc = f(init_damp), f(min_damp), f(max_damp), f(damp_dec), f(damp_inc)
    jac_thresh, min_grad, min_param_inc = f(jac_thresh), f(min_grad), f(min_param_inc)
    R, t, scale = np.asarray(R, f), np.asarray(t, f), f(scale)
    prev_error, curr_error = f(0), f(1)
    damp, it = f(init_damp), 0
    trace = []
    cand_err = curr_error
    while True:
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = abs(curr_error - prev_error) / prev_error
        if ratio > jac_thresh:
            A, b, e = jac_fn(R, t, scale)
            AtA, Atb = np.asarray(A, f), np.asarray(b, f).reshape(-1)
            if it == 0:
                curr_error = f(e)
            update_jac = True
        else:
            update_jac = False
        it += 1
        diag = np.diag(np.diag(AtA))
        sol = np.linalg.solve((AtA + damp * diag).astype(np.float64), Atb.astype(np.float64)).astype(f)
        rotvec = rotation_to_angle_axis(R).astype(f)
        x = np.concatenate([t.reshape(-1), rotvec, [scale]])
        if np.abs(Atb).max() < min_grad or (sol / (np.abs(x) + f(1e-8))).max() < min_param_inc:
            break
        while True:
            Rc, tc = retract(R, t, sol[:6])
            Rc, tc, sc = Rc.astype(f), tc.astype(f), f(scale + sol[6])
            cand_err = f(err_fn(Rc, tc, sc))
            if cand_err < curr_error:
                break
            elif damp < max_damp:
                damp = f(min(max(min_damp, damp * damp_inc), max_damp))
                sol = np.linalg.solve((AtA + damp * diag).astype(np.float64), Atb.astype(np.float64)).astype(f)
            else:
                break
        accepted = not (cand_err >= curr_error and damp >= max_damp)
        trace.append((it, float(damp), bool(accepted), float(cand_err)))
        if not accepted:
            break
        R, t, scale = Rc, tc, sc
        if update_jac:
            prev_error = curr_error
        curr_error = cand_err
        damp = f(min(max(min_damp, damp / damp_dec), max_damp))
        if it >= max_iters:
            break
    return R, t, float(scale), float(curr_error), trace


def tracker_lm(jac_fn, err_fn, R, t, init_damp=1e-4, min_damp=1e-6, max_damp=1e-2, damp_dec=10.0, damp_inc=100.0,
               max_iters=40, jac_thresh=1e-2, min_grad=1e-8, min_param_inc=1e-8):
    """CameraTracker::TrackNewFrame LM loop (core/system/camera_tracker.cpp:1156-1279) over
    user-supplied jac_fn(R,t)->(AtA,Atb,err) / err_fn(R,t)->err.  fp32 state like the reference.
    Returns (R, t, error, trace) where trace records (iter, damp, accepted, error)."""
    f = np.float32
    # float options like the reference's (see tracker_lm7)
    init_damp, min_damp, max_damp, damp_dec, damp_inc = f(init_damp), f(min_damp), f(max_damp), f(damp_dec), f(damp_inc)
    jac_thresh, min_grad, min_param_inc = f(jac_thresh), f(min_grad), f(min_param_inc)
    R, t = np.asarray(R, f), np.asarray(t, f)
    prev_error, curr_error = f(0), f(1)
    damp, it = f(init_damp), 0
    AtA = Atb = None
    trace = []
    cand_err = curr_error
    while True:
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = abs(curr_error - prev_error) / prev_error
        if ratio > jac_thresh:
            A, b, e = jac_fn(R, t)
            AtA, Atb = np.asarray(A, f), np.asarray(b, f).reshape(-1)
            if it == 0:
                curr_error = f(e)
            update_jac = True
        else:
            update_jac = False
        it += 1
        diag = np.diag(np.diag(AtA))
        sol = np.linalg.solve((AtA + damp * diag).astype(np.float64), Atb.astype(np.float64)).astype(f)
        rotvec = rotation_to_angle_axis(R).astype(f)
        max_grad = np.abs(Atb).max()
        max_inc = (sol / (np.abs(np.concatenate([t.reshape(-1), rotvec])) + f(1e-8))).max()
        if max_grad < min_grad or max_inc < min_param_inc:
            break
        while True:
            Rc, tc = retract(R, t, sol)
            Rc, tc = Rc.astype(f), tc.astype(f)
            cand_err = f(err_fn(Rc, tc))
            if cand_err < curr_error:
                break
            elif damp < max_damp:
                damp = f(min(max(min_damp, damp * damp_inc), max_damp))
                sol = np.linalg.solve((AtA + damp * diag).astype(np.float64), Atb.astype(np.float64)).astype(f)
            else:
                break
        accepted = not (cand_err >= curr_error and damp >= max_damp)
        trace.append((it, float(damp), bool(accepted), float(cand_err)))
        if not accepted:
            break
        R, t = Rc, tc
        if update_jac:
            prev_error = curr_error
        curr_error = cand_err
        damp = f(min(max(min_damp, damp / damp_dec), max_damp))
        if it >= max_iters:
            break
    return R, t, float(curr_error), trace

"""Host-side mirror of the reference operator interface over the C ABI.

Function names and argument meaning follow cuda/*_factor_kernels.h of the reference
(`photometric_jac_error_calculate`, `geometric_error_calculate`, ...); tensors become numpy arrays
and the per-frame tensors (`kf->feat_map_pyramid`, `kf->dpt_jac_code`, ...) travel inside a
DeviceKeyframe, which is the once-per-keyframe device copy of a frames.Keyframe.  Every call goes
through libsage_ba.so -- there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import capi
from .frames import Keyframe

F32 = np.float32


def _f(a):
    return np.ascontiguousarray(a, dtype=F32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SageError(RuntimeError):
    pass


class Context:
    """sage_ba_context: one per host thread / CUDA stream (the reference is re-entrant from 4 threads)."""

    def __init__(self, device=0, stream=None):
        self.lib = capi.load()
        h = C.c_void_p()
        if self.lib.sage_ba_create(C.byref(h), int(device), C.c_void_p(stream) if stream else None) != 0:
            raise SageError("sage_ba_create failed: no usable CUDA device (the factor kernels have no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise SageError(self.lib.sage_ba_last_error(self.h).decode())

    def synchronize(self):
        self.check(self.lib.sage_ba_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.sage_ba_launch_count(self.h))

    @property
    def stream_handle(self):
        """cudaStream_t (as int) every call of this context is enqueued on."""
        return int(self.lib.sage_ba_stream(self.h) or 0)

    def close(self):
        if self.h:
            self.lib.sage_ba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceKeyframe:
    """Device-resident, re-laid-out copy of a Keyframe (sage_ba_keyframe)."""

    def __init__(self, ctx: Context, kf: Keyframe, with_depth=True, build_pyramid_on_device=False):
        """build_pyramid_on_device: hand over only the level-0 feature map [F,H,W] (the feature net's output) and let
        the library build the masked Gaussian pyramid + gradients (Mapper::GenerateGaussianPyramidWithGrad)."""
        self.ctx, self.kf = ctx, kf
        H, W = kf.video_mask.shape
        cam = kf.camera_pyramid[0]
        d = capi.KeyframeDesc()
        d.memory = capi.HOST
        d.height, d.width, d.levels = H, W, len(kf.camera_pyramid)
        d.feat_channels = kf.feat_map_pyramid.shape[0]
        d.code_size = kf.dpt_jac_code.shape[1]
        d.camera = capi.Camera(*[float(x) for x in cam])
        self._keep = []

        def hold(a, dtype=F32):
            a = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        if build_pyramid_on_device:
            d.feat_map = hold(kf.feat_map_pyramid[:, :H * W])
        else:
            d.feat_map_pyramid = hold(kf.feat_map_pyramid)
            d.feat_map_grad_pyramid = hold(kf.feat_map_grad_pyramid) if kf.feat_map_grad_pyramid is not None else None
        d.video_mask = hold(kf.video_mask)
        if with_depth:
            d.dpt_map_bias = hold(kf.dpt_map_bias)
            jac = kf.dpt_jac_code
            # hand over the strided view as the reference does: base pointer + element strides
            base = jac.base if jac.base is not None and jac.base.dtype == F32 else None
            if base is not None and base.flags["C_CONTIGUOUS"] and jac.ctypes.data == base.ctypes.data:
                self._keep.append(base)
                d.dpt_jac_code = base.ctypes.data_as(C.c_void_p)
                d.jac_stride_row = jac.strides[0] // 4
                d.jac_stride_col = jac.strides[1] // 4
            else:
                d.dpt_jac_code = hold(jac)
                d.jac_stride_row, d.jac_stride_col = jac.shape[1], 1
            d.sampled_locations_1d = hold(kf.sampled_locations_1d, np.int64)
        d.sampled_locations_homo = hold(kf.sampled_locations_homo)
        d.num_samples = len(kf.sampled_locations_homo)
        h = C.c_void_p()
        ctx.check(ctx.lib.sage_ba_keyframe_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self._keep = []  # create() is synchronous: host staging no longer needed
        self.H, self.W, self.L = H, W, d.levels
        self.F, self.C, self.N = d.feat_channels, d.code_size, d.num_samples

    def cameras(self):
        cams = (capi.Camera * self.L)()
        offs = (C.c_int * self.L)()
        self.ctx.lib.sage_ba_keyframe_cameras(self.h, cams, offs)
        return np.array([[c.fx, c.fy, c.u0, c.v0, c.width, c.height] for c in cams], F32), np.array(list(offs), np.int32)

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sage_ba_keyframe_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------
# reference operator API (cuda/photometric_factor_kernels.h, geometric_..., reprojection_...)
# --------------------------------------------------------------------------------------------------
def photometric_jac_error_calculate(ctx, kf0, kf1, rotation10, translation10, rotation0, translation0, rotation1,
                                    translation1, code_0, scale_0, eps, weights):
    D = 13 + kf0.C
    AtA, Atb = np.zeros((D, D), F32), np.zeros((D,), F32)
    err, inl = C.c_float(0), C.c_float(0)
    a = [_f(x) for x in (rotation10, translation10, rotation0, translation0, rotation1, translation1, code_0)]
    w = _f(weights)
    ctx.check(ctx.lib.sage_ba_photometric_jac_error(ctx.h, kf0.h, kf1.h, *[_p(x) for x in a], float(scale_0), float(eps),
                                                    _p(w), _p(AtA), _p(Atb), C.byref(err), C.byref(inl)))
    return AtA, Atb, err.value, inl.value


def photometric_error_calculate(ctx, kf0, kf1, rotation, translation, code_0, scale_0, eps, weights):
    err, inl = C.c_float(0), C.c_float(0)
    R, t, c, w = _f(rotation), _f(translation), _f(code_0), _f(weights)
    ctx.check(ctx.lib.sage_ba_photometric_error(ctx.h, kf0.h, kf1.h, _p(R), _p(t), _p(c), float(scale_0), float(eps), _p(w),
                                                C.byref(err), C.byref(inl)))
    return err.value, inl.value


def geometric_jac_error_calculate(ctx, kf0, kf1, rotation10, translation10, rotation0, translation0, rotation1, translation1,
                                  code_0, code_1, scale_0, scale_1, eps, loss_param, weight):
    D = 14 + 2 * kf0.C
    AtA, Atb = np.zeros((D, D), F32), np.zeros((D,), F32)
    err, inl = C.c_float(0), C.c_float(0)
    a = [_f(x) for x in (rotation10, translation10, rotation0, translation0, rotation1, translation1, code_0, code_1)]
    ctx.check(ctx.lib.sage_ba_geometric_jac_error(ctx.h, kf0.h, kf1.h, *[_p(x) for x in a], float(scale_0), float(scale_1),
                                                  float(eps), float(loss_param), float(weight), _p(AtA), _p(Atb),
                                                  C.byref(err), C.byref(inl)))
    return AtA, Atb, err.value, inl.value


def geometric_error_calculate(ctx, kf0, kf1, rotation, translation, code_0, code_1, scale_0, scale_1, eps, loss_param, weight):
    err, inl = C.c_float(0), C.c_float(0)
    a = [_f(x) for x in (rotation, translation, code_0, code_1)]
    ctx.check(ctx.lib.sage_ba_geometric_error(ctx.h, kf0.h, kf1.h, *[_p(x) for x in a], float(scale_0), float(scale_1),
                                              float(eps), float(loss_param), float(weight), C.byref(err), C.byref(inl)))
    return err.value, inl.value


def reprojection_jac_error_calculate(ctx, kf0, rotation10, translation10, rotation0, translation0, rotation1, translation1,
                                     code_0, scale_0, matched_locations_1d_0, matched_locations_homo_0,
                                     matched_locations_2d_1, eps, loss_param, weight):
    D = 13 + kf0.C
    AtA, Atb = np.zeros((D, D), F32), np.zeros((D,), F32)
    err, inl = C.c_float(0), C.c_float(0)
    a = [_f(x) for x in (rotation10, translation10, rotation0, translation0, rotation1, translation1, code_0)]
    loc = np.ascontiguousarray(matched_locations_1d_0, np.int32)
    homo, m2d = _f(matched_locations_homo_0), _f(matched_locations_2d_1)
    ctx.check(ctx.lib.sage_ba_reprojection_jac_error(ctx.h, kf0.h, *[_p(x) for x in a], float(scale_0), _p(loc), _p(homo),
                                                     _p(m2d), len(loc), float(eps), float(loss_param), float(weight),
                                                     _p(AtA), _p(Atb), C.byref(err), C.byref(inl)))
    return AtA, Atb, err.value, inl.value


def reprojection_error_calculate(ctx, kf0, rotation10, translation10, code_0, scale_0, matched_locations_1d_0,
                                 matched_locations_homo_0, matched_locations_2d_1, eps, loss_param, weight):
    err, inl = C.c_float(0), C.c_float(0)
    R, t, c = _f(rotation10), _f(translation10), _f(code_0)
    loc = np.ascontiguousarray(matched_locations_1d_0, np.int32)
    homo, m2d = _f(matched_locations_homo_0), _f(matched_locations_2d_1)
    ctx.check(ctx.lib.sage_ba_reprojection_error(ctx.h, kf0.h, _p(R), _p(t), _p(c), float(scale_0), _p(loc), _p(homo), _p(m2d),
                                                 len(loc), float(eps), float(loss_param), float(weight), C.byref(err),
                                                 C.byref(inl)))
    return err.value, inl.value


def tracker_reproj_jac_error_calculate(ctx, camera, rotation, translation, sampled_dpts_0, sampled_locations_homo_0,
                                       matched_locations_2d_1, eps, loss_param, weight):
    AtA, Atb = np.zeros((6, 6), F32), np.zeros((6,), F32)
    err, inl = C.c_float(0), C.c_float(0)
    cam = capi.Camera(*[float(x) for x in camera])
    R, t, d, h, m = [_f(x) for x in (rotation, translation, sampled_dpts_0, sampled_locations_homo_0, matched_locations_2d_1)]
    ctx.check(ctx.lib.sage_ba_tracker_reproj_jac_error(ctx.h, C.byref(cam), _p(R), _p(t), _p(d), _p(h), _p(m), len(d), float(eps),
                                                       float(loss_param), float(weight), _p(AtA), _p(Atb), C.byref(err),
                                                       C.byref(inl)))
    return AtA, Atb, err.value, inl.value


def tracker_reproj_error_calculate(ctx, camera, rotation, translation, sampled_dpts_0, sampled_locations_homo_0,
                                   matched_locations_2d_1, eps, loss_param, weight):
    err, inl = C.c_float(0), C.c_float(0)
    cam = capi.Camera(*[float(x) for x in camera])
    R, t, d, h, m = [_f(x) for x in (rotation, translation, sampled_dpts_0, sampled_locations_homo_0, matched_locations_2d_1)]
    ctx.check(ctx.lib.sage_ba_tracker_reproj_error(ctx.h, C.byref(cam), _p(R), _p(t), _p(d), _p(h), _p(m), len(d), float(eps),
                                                   float(loss_param), float(weight), C.byref(err), C.byref(inl)))
    return err.value, inl.value


class TrackerSamples:
    """Device tensors the tracker pre-samples once per reference keyframe
    (camera_tracker.cpp:1086-1123): sampled_dpts_0 [N], sampled_locations_homo_0 [N,3],
    sampled_features_0 [L,N,F].  Device memory is owned through torch."""

    def __init__(self, ctx, kf0: DeviceKeyframe, code_0, scale_0):
        import torch

        dev = torch.device("cuda", ctx.device)
        self.dpts = torch.empty(kf0.N, dtype=torch.float32, device=dev)
        self.homo = torch.empty(kf0.N, 3, dtype=torch.float32, device=dev)
        self.feats = torch.empty(kf0.L, kf0.N, kf0.F, dtype=torch.float32, device=dev)
        self.N = kf0.N
        c = _f(code_0)
        ctx.check(ctx.lib.sage_ba_tracker_presample(ctx.h, kf0.h, _p(c), float(scale_0), C.c_void_p(self.dpts.data_ptr()),
                                                    C.c_void_p(self.homo.data_ptr()), C.c_void_p(self.feats.data_ptr())))


def tracker_photo_jac_error_calculate(ctx, frame1, rotation, translation, samples: TrackerSamples, eps, weights, scale_0=None):
    """tracker_photo_jac_error_calculate (scale_0 None, 6x6) or ..._with_scale (7x7)."""
    D = 7 if scale_0 is not None else 6
    AtA, Atb = np.zeros((D, D), F32), np.zeros((D,), F32)
    err, inl = C.c_float(0), C.c_float(0)
    R, t, w = _f(rotation), _f(translation), _f(weights)
    ctx.check(ctx.lib.sage_ba_tracker_photo_jac_error(
        ctx.h, frame1.h, _p(R), _p(t), C.c_void_p(samples.dpts.data_ptr()), C.c_void_p(samples.homo.data_ptr()),
        C.c_void_p(samples.feats.data_ptr()), samples.N, int(scale_0 is not None), float(scale_0 or 0.0), float(eps), _p(w),
        _p(AtA), _p(Atb), C.byref(err), C.byref(inl)))
    return AtA, Atb, err.value, inl.value


def tracker_photo_error_calculate(ctx, frame1, rotation, translation, samples: TrackerSamples, eps, weights):
    err, inl = C.c_float(0), C.c_float(0)
    R, t, w = _f(rotation), _f(translation), _f(weights)
    ctx.check(ctx.lib.sage_ba_tracker_photo_error(
        ctx.h, frame1.h, _p(R), _p(t), C.c_void_p(samples.dpts.data_ptr()), C.c_void_p(samples.homo.data_ptr()),
        C.c_void_p(samples.feats.data_ptr()), samples.N, float(eps), _p(w), C.byref(err), C.byref(inl)))
    return err.value, inl.value


def track_new_frame(ctx, kf0, frame1, code_0, scale_0, rotation, translation, photo_weights, dpt_eps=1e-4, max_num_iters=40,
                    init_damp=1e-4, min_damp=1e-6, max_damp=1e-2, damp_dec_factor=10.0, damp_inc_factor=100.0,
                    jac_update_err_inc_threshold=1e-2, min_grad_thresh=1e-8, min_param_inc_thresh=1e-8, matches=None,
                    reproj_loss_param=1.0, reproj_weight=0.0, use_photo=True):
    """CameraTracker::TrackNewFrame (camera_tracker.cpp:1034-1310). matches = (dpts [M], homo [M,3], 2d [M,2]) or None."""
    cfg = capi.TrackerConfig()
    cfg.max_num_iters = max_num_iters
    cfg.init_damp, cfg.min_damp, cfg.max_damp = init_damp, min_damp, max_damp
    cfg.damp_dec_factor, cfg.damp_inc_factor = damp_dec_factor, damp_inc_factor
    cfg.jac_update_err_inc_threshold = jac_update_err_inc_threshold
    cfg.min_grad_thresh, cfg.min_param_inc_thresh, cfg.dpt_eps = min_grad_thresh, min_param_inc_thresh, dpt_eps
    for i, w in enumerate(photo_weights):
        cfg.photo_weights[i] = float(w)
    cfg.use_photo, cfg.use_reproj = int(use_photo), int(matches is not None)
    cfg.reproj_loss_param, cfg.reproj_weight = reproj_loss_param, reproj_weight
    R, t, c = _f(rotation).copy(), _f(translation).copy(), _f(code_0)
    md = mh = m2 = None
    M = 0
    if matches is not None:
        md, mh, m2 = [_f(x) for x in matches]
        M = len(md)
    rep = capi.TrackerReport()
    ctx.check(ctx.lib.sage_ba_track_new_frame(ctx.h, kf0.h, frame1.h, _p(c), float(scale_0), C.byref(cfg), _p(R), _p(t), _p(md),
                                              _p(mh), _p(m2), M, C.byref(rep)))
    return R, t, {"iterations": rep.iterations, "jacobian_evals": rep.jacobian_evals, "error_evals": rep.error_evals,
                  "final_error": rep.final_error, "final_damp": rep.final_damp}


def tracker_match_geom_jac_error_calculate(ctx, rotation, translation, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0,
                                           matched_locations_homo_1, loss_param, weight, scale_0=None):
    """tracker_match_geom_jac_error_calculate (6x6) or ..._with_scale (7x7) (cuda/match_geometry_factor_kernels.h)."""
    D = 7 if scale_0 is not None else 6
    AtA, Atb = np.zeros((D, D), F32), np.zeros((D,), F32)
    err = C.c_float(0)
    R, t, d0, d1, h0, h1 = [_f(x) for x in (rotation, translation, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0,
                                            matched_locations_homo_1)]
    ctx.check(ctx.lib.sage_ba_tracker_match_geom_jac_error(ctx.h, _p(R), _p(t), _p(d0), _p(d1), _p(h0), _p(h1), len(d0),
                                                           int(scale_0 is not None), float(scale_0 or 0.0), float(loss_param),
                                                           float(weight), _p(AtA), _p(Atb), C.byref(err)))
    return AtA, Atb, err.value


def tracker_match_geom_error_calculate(ctx, rotation, translation, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0,
                                       matched_locations_homo_1, loss_param, weight):
    err = C.c_float(0)
    R, t, d0, d1, h0, h1 = [_f(x) for x in (rotation, translation, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0,
                                            matched_locations_homo_1)]
    ctx.check(ctx.lib.sage_ba_tracker_match_geom_error(ctx.h, _p(R), _p(t), _p(d0), _p(d1), _p(h0), _p(h1), len(d0),
                                                       float(loss_param), float(weight), C.byref(err)))
    return err.value


LOSS_TYPES = {"fair": 0, "L2": 1, "huber": 2, "unbiased": 3}  # robust_loss_type strings of match_geometry_factor_kernels.h


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def match_geometry_jac_error_calculate(ctx, kf0, kf1, rotation10, translation10, rotation0, translation0, rotation1, translation1,
                                       code_0, code_1, sampled_locations_homo_0, matched_locations_homo_1, sampled_locations_1d_0,
                                       matched_locations_1d_1, scale_0, scale_1, loss_param, weight, robust_loss_type="fair"):
    """match_geometry_jac_error_calculate<CS> (cuda/match_geometry_factor_kernels.h:41-56); the per-frame bias / basis tensors of the
    reference signature live in the keyframe handles."""
    D = 14 + 2 * kf0.C
    AtA, Atb = np.zeros((D, D), F32), np.zeros((D,), F32)
    err = C.c_float(0)
    a = [_f(x) for x in (rotation10, translation10, rotation0, translation0, rotation1, translation1, code_0, code_1)]
    h0, h1 = _f(sampled_locations_homo_0), _f(matched_locations_homo_1)
    l0, l1 = _i32(sampled_locations_1d_0), _i32(matched_locations_1d_1)
    ctx.check(ctx.lib.sage_ba_match_geometry_jac_error(ctx.h, kf0.h, kf1.h, *[_p(x) for x in a], float(scale_0), float(scale_1), _p(l0),
                                                       _p(l1), _p(h0), _p(h1), len(l0), float(loss_param), float(weight),
                                                       LOSS_TYPES[robust_loss_type], _p(AtA), _p(Atb), C.byref(err)))
    return AtA, Atb, err.value


def match_geometry_error_calculate(ctx, kf0, kf1, rotation, translation, code_0, code_1, sampled_locations_homo_0,
                                   matched_locations_homo_1, sampled_locations_1d_0, matched_locations_1d_1, scale_0, scale_1,
                                   loss_param, weight, robust_loss_type="fair"):
    err = C.c_float(0)
    a = [_f(x) for x in (rotation, translation, code_0, code_1)]
    h0, h1 = _f(sampled_locations_homo_0), _f(matched_locations_homo_1)
    l0, l1 = _i32(sampled_locations_1d_0), _i32(matched_locations_1d_1)
    ctx.check(ctx.lib.sage_ba_match_geometry_error(ctx.h, kf0.h, kf1.h, *[_p(x) for x in a], float(scale_0), float(scale_1), _p(l0),
                                                   _p(l1), _p(h0), _p(h1), len(l0), float(loss_param), float(weight),
                                                   LOSS_TYPES[robust_loss_type], C.byref(err)))
    return err.value


def loop_mg_jac_error_calculate(ctx, rotation10, translation10, rotation0, translation0, rotation1, translation1,
                                sampled_unscaled_dpts_0, matched_unscaled_dpts_1, sampled_locations_homo_0, matched_locations_homo_1,
                                scale_0, scale_1, loss_param, weight):
    """loop_mg_jac_error_calculate (cuda/match_geometry_factor_kernels.h:67-77): 14x14, order [pose0 pose1 scale0 scale1]."""
    AtA, Atb = np.zeros((14, 14), F32), np.zeros((14,), F32)
    err = C.c_float(0)
    a = [_f(x) for x in (rotation10, translation10, rotation0, translation0, rotation1, translation1, sampled_unscaled_dpts_0,
                         matched_unscaled_dpts_1, sampled_locations_homo_0, matched_locations_homo_1)]
    ctx.check(ctx.lib.sage_ba_loop_mg_jac_error(ctx.h, *[_p(x) for x in a], len(a[6]), float(scale_0), float(scale_1), float(loss_param),
                                                float(weight), _p(AtA), _p(Atb), C.byref(err)))
    return AtA, Atb, err.value


def loop_mg_error_calculate(ctx, rotation, translation, sampled_unscaled_dpts_0, matched_unscaled_dpts_1, sampled_locations_homo_0,
                            matched_locations_homo_1, scale_0, scale_1, loss_param, weight):
    err = C.c_float(0)
    a = [_f(x) for x in (rotation, translation, sampled_unscaled_dpts_0, matched_unscaled_dpts_1, sampled_locations_homo_0,
                         matched_locations_homo_1)]
    ctx.check(ctx.lib.sage_ba_loop_mg_error(ctx.h, *[_p(x) for x in a], len(a[2]), float(scale_0), float(scale_1), float(loss_param),
                                            float(weight), C.byref(err)))
    return err.value


def cycle_feature_matching(ctx, feat_desc_0, feat_desc_1, keypoint_locations_1d, cyc_consis_thresh, device_ptrs=None, timing=False):
    """Dense descriptor cycle-matching of the factor constructors / FeatureMatchingGeo
    (core/gtsam/reprojection_factor.cpp:57-92, core/system/camera_tracker.cpp:798-834).

    feat_desc_{0,1}: [C, H, W] (or [1, C, H, W]) float32, channel-major like Frame::feat_desc.
    keypoint_locations_1d: [K] = valid_locations_1d[keypoint_indexes].
    device_ptrs=(ptr0, ptr1, C, H, W): the maps already live on the device (the arrays are then ignored).
    Returns dict(raw_matched_locations_1d_1 [K], cyc_matched_locations_1d_0 [K], inlier_within_keypoint_indexes [M],
    matched_locations_1d_1 [M], matched_locations_2d_1 [M,2]) -- names as in the reference -- plus kernel_ms with timing=True."""
    kp = np.ascontiguousarray(keypoint_locations_1d, dtype=np.int64)
    K = len(kp)
    if device_ptrs is None:
        d0 = _f(feat_desc_0).reshape((-1,) + tuple(np.shape(feat_desc_0)[-2:]))
        d1 = _f(feat_desc_1).reshape(d0.shape)
        Cd, H, W = d0.shape
        p0, p1, mem = _p(d0), _p(d1), capi.HOST
    else:
        p0, p1, Cd, H, W = device_ptrs
        p0, p1, mem = C.c_void_p(p0), C.c_void_p(p1), capi.DEVICE
    raw, cyc, sel = np.zeros(K, np.int32), np.zeros(K, np.int32), np.zeros(K, np.int32)
    m, ms = C.c_int(0), C.c_float(0)
    ctx.check(ctx.lib.sage_ba_cycle_match(ctx.h, mem, p0, p1, int(Cd), int(H), int(W), _p(kp), K, float(cyc_consis_thresh), _p(raw),
                                          _p(cyc), _p(sel), C.byref(m), C.byref(ms) if timing else None))
    sel = sel[:m.value].astype(np.int64)
    loc1 = raw[sel]
    out = {"raw_matched_locations_1d_1": raw, "cyc_matched_locations_1d_0": cyc, "inlier_within_keypoint_indexes": sel,
           "matched_locations_1d_1": loc1,
           "matched_locations_2d_1": np.stack([np.fmod(loc1.astype(F32), F32(W)), np.floor(loc1.astype(F32) / F32(W))], 1)}
    if timing:
        out["kernel_ms"] = ms.value
    return out


def track_frame(ctx, frame0, kf1, code_0, rotation, translation, scale, photo_weights, dpt_eps=1e-4, max_num_iters=40, init_damp=1e-4,
                min_damp=1e-6, max_damp=1e-2, damp_dec_factor=10.0, damp_inc_factor=100.0, jac_update_err_inc_threshold=1e-2,
                min_grad_thresh=1e-8, min_param_inc_thresh=1e-8, matches=None, match_geom_loss_param=1.0, match_geom_weight=0.0,
                use_photo=True):
    """CameraTracker::TrackFrame (camera_tracker.cpp:1312-1672): 7-DoF LM on relative pose + depth scale of frame0.
    matches = (unscaled_dpts_0 [M], homo_0 [M,3], dpts_1 [M], homo_1 [M,3]) or None.  Returns (R, t, scale, report)."""
    cfg = capi.TrackerConfig()
    cfg.max_num_iters = max_num_iters
    cfg.init_damp, cfg.min_damp, cfg.max_damp = init_damp, min_damp, max_damp
    cfg.damp_dec_factor, cfg.damp_inc_factor = damp_dec_factor, damp_inc_factor
    cfg.jac_update_err_inc_threshold = jac_update_err_inc_threshold
    cfg.min_grad_thresh, cfg.min_param_inc_thresh, cfg.dpt_eps = min_grad_thresh, min_param_inc_thresh, dpt_eps
    for i, w in enumerate(photo_weights):
        cfg.photo_weights[i] = float(w)
    cfg.use_photo, cfg.use_match_geom = int(use_photo), int(matches is not None)
    cfg.match_geom_loss_param, cfg.match_geom_weight = match_geom_loss_param, match_geom_weight
    R, t, c = _f(rotation).copy(), _f(translation).copy(), _f(code_0)
    s = C.c_float(float(scale))
    arrs = [None] * 4
    M = 0
    if matches is not None:
        arrs = [_f(x) for x in matches]
        M = len(arrs[0])
    rep = capi.TrackerReport()
    rc = ctx.lib.sage_ba_track_frame(ctx.h, frame0.h, kf1.h, _p(c), C.byref(cfg), _p(R), _p(t), C.byref(s), _p(arrs[0]), _p(arrs[1]),
                                     _p(arrs[2]), _p(arrs[3]), M, C.byref(rep))
    if rc == 1:
        ctx.check(rc)
    return R, t, s.value, {"iterations": rep.iterations, "jacobian_evals": rep.jacobian_evals, "error_evals": rep.error_evals,
                           "final_error": rep.final_error, "final_damp": rep.final_damp, "no_overlap": rc == 2}

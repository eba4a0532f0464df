"""Batched stand-in for Mapper::MappingStep / Mapper::UpdateMap.

The reference keeps its factor graph in GTSAM ISAM2: EnqueueKeyframe / EnqueueLink turn into work items, Bookkeeping turns
work items into factors, MappingStep calls isam_graph_->update(...) and UpdateMap copies the estimate back into the
keyframes (core/mapping/mapper.cpp:150-198, :300-445, :469-612, :1141-1180).  This adapter keeps the same calls and
the same factor set per link, but the solve is the device-side batched LM of libsage_ba (all factors of the window are
re-linearised every iteration; no incremental elimination), and the write-back runs UpdateDepth on the device.

Not reproduced (out of scope): work-item bookkeeping/removal, marginalisation, loop-closure links,
and TEASER++ filtering of the descriptor matches (reprojection_factor.cpp:136-186) -- cycle-consistent matches are used as is.
The keypoint draw is the reference's std::shuffle with std::mt19937(kf.id * fr.id) (frames.std_shuffle).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from . import capi, factors, ops
from .frames import Keyframe
from .local_ba import LocalBA

F32 = np.float32


@dataclass
class MapperOptions:
    """The DeepFactorsOptions fields MappingStep's factors read (core/deepfactors_options.h, configs/slam_run.flags:96-106)."""
    use_photometric: bool = True
    use_reprojection: bool = True
    use_geometric: bool = True
    photo_factor_weights: Tuple[float, ...] = (10.0, 9.0, 8.0, 7.0)
    geo_factor_weight: float = 0.1
    geo_loss_param_factor: float = 0.03
    reproj_factor_weight: float = 0.1
    reproj_loss_param_factor: float = 0.03
    code_factor_weight: float = 1e-3
    init_scale_prior_weight: float = 1e-2
    dpt_eps: float = 1e-4
    desc_num_keypoints: int = 512
    desc_cyc_consis_thresh: float = 2.0
    factor_iters: int = 10  # LM iterations per mapping_step


@dataclass
class Map:
    """df::Map (core/mapping/keyframe_map.h:93-120): keyframes by id + directed links."""
    keyframes: Dict[int, Keyframe] = field(default_factory=dict)
    links: List[Tuple[int, int]] = field(default_factory=list)


def avg_squared_dpt_bias(kf: Keyframe) -> float:
    """Frame::avg_squared_dpt_bias as Mapper::BuildKeyframe sets it (mapper.cpp:1376-1378): the MASKED mean
    sum((dpt_map_bias * mask)^2) / sum(mask) in fp32 -- with an endoscope mask this differs from the plain mean over all pixels,
    and it scales the robust loss of every geometric factor."""
    m = np.asarray(kf.video_mask, F32).reshape(-1)
    b = np.asarray(kf.dpt_map_bias, F32).reshape(-1)
    return float(np.sum(np.square(b * m, dtype=F32), dtype=F32) / np.sum(m, dtype=F32))


class BatchedMapper:
    def __init__(self, ctx: ops.Context, opts: MapperOptions = None):
        self.ctx = ctx
        self.opts = opts or MapperOptions()
        self.map = Map()
        self._dev: Dict[int, ops.DeviceKeyframe] = {}
        self._order: List[int] = []                      # keyframe ids in insertion order = problem index
        self._factors: List[tuple] = []                  # ("photo"|"geo"|"reproj", id0, id1, payload)
        self._code_priors: List[int] = []
        self._scale_priors: List[Tuple[int, float]] = []
        self._fixed: List[int] = []
        self.last_report = None
        self.match_stats: Dict[Tuple[int, int], Tuple[int, int]] = {}

    # ------------------------------------------------------------------------------------------ keyframes / links
    def _add_kf(self, kf: Keyframe):
        assert kf.id not in self.map.keyframes, "keyframe id already in the map"
        self.map.keyframes[kf.id] = kf
        self._dev[kf.id] = ops.DeviceKeyframe(self.ctx, kf)
        self._order.append(kf.id)

    def init_one_frame(self, kf: Keyframe):
        """Mapper::InitOneFrame (mapper.cpp:150-198): the first keyframe is the gauge -- depth normalised by its median,
        pose held, scale prior, code prior."""
        valid = kf.dpt_map.reshape(-1)[kf.sampled_locations_1d]
        median = float(np.sort(valid)[(len(valid) - 1) // 2])  # torch::median returns the lower middle element
        kf.dpt_scale = float(F32(kf.dpt_scale) / F32(median))
        kf.dpt_map_stored = None
        self._add_kf(kf)
        self._fixed.append(kf.id)
        self._scale_priors.append((kf.id, kf.dpt_scale))
        self._code_priors.append(kf.id)

    def enqueue_keyframe(self, kf: Keyframe, conns: List[int]):
        """Mapper::EnqueueKeyframe (mapper.cpp:300-380): code prior + the enabled factor kinds both ways per connection."""
        kf.temporal_connections = list(conns)
        self._add_kf(kf)
        self._code_priors.append(kf.id)
        for back in conns:
            self._link(kf.id, back)

    def enqueue_link(self, id0: int, id1: int, photo=True, rep=True, geo=True):
        """Mapper::EnqueueLink (mapper.cpp:395-445)."""
        self._link(id0, id1, photo, rep, geo)

    def _link(self, a, b, photo=True, rep=True, geo=True):
        o = self.opts
        self.map.links.append((a, b))
        for i, j in ((a, b), (b, a)):
            if o.use_photometric and photo:
                self._factors.append(("photo", i, j, None))
            if o.use_reprojection and rep:
                m = self._match(i, j)
                if m is not None:
                    self._factors.append(("reproj", i, j, m))
            if o.use_geometric and geo:
                kf = self.map.keyframes[a]
                loss = o.geo_loss_param_factor * avg_squared_dpt_bias(kf)
                self._factors.append(("geo", i, j, loss))

    def _match(self, i, j):
        """ReprojectionFactor's constructor up to the TEASER step (reprojection_factor.cpp:36-112)."""
        kf, fr = self.map.keyframes[i], self.map.keyframes[j]
        if kf.feat_desc is None or fr.feat_desc is None:
            return None
        r = factors.cycle_matches(self.ctx, kf, fr, self.opts.desc_num_keypoints, self.opts.desc_cyc_consis_thresh)
        K = min(self.opts.desc_num_keypoints, len(kf.sampled_locations_1d))
        self.match_stats[(i, j)] = (0 if r is None else len(r["matched_locations_1d_0"]), K)
        if r is None:
            return None
        return (r["matched_locations_1d_0"], r["matched_locations_homo_0"], r["matched_locations_2d_1"])

    # ------------------------------------------------------------------------------------------ MappingStep
    def mapping_step(self, iters=None):
        """Mapper::MappingStep (mapper.cpp:469-612): optimise every variable touched by the current factors, then UpdateMap."""
        o = self.opts
        ids = self._order
        pos = {k: n for n, k in enumerate(ids)}
        kfs = [self.map.keyframes[k] for k in ids]
        ba = LocalBA(self.ctx, [self._dev[k] for k in ids])
        W = kfs[0].video_mask.shape[1]
        for kind, i, j, payload in self._factors:
            if kind == "photo":
                ba.add_photometric(pos[i], pos[j], o.photo_factor_weights[:ba.L])
            elif kind == "geo":
                ba.add_geometric(pos[i], pos[j], payload, o.geo_factor_weight)
            else:
                loc, homo, uv = payload
                ba.add_reprojection(pos[i], pos[j], loc, homo, uv, o.reproj_loss_param_factor * W * W, o.reproj_factor_weight)
        for k in self._code_priors:
            ba.add_code_prior(pos[k], o.code_factor_weight)
        for k, s in self._scale_priors:
            ba.add_scale_prior(pos[k], s, o.init_scale_prior_weight)
        for k in self._fixed:
            ba.fix(pos[k], pose=True, scale=False)
        ba.set_state([kf.pose_wk for kf in kfs], np.stack([kf.code for kf in kfs]), [kf.dpt_scale for kf in kfs], eps=o.dpt_eps)
        self.last_report = ba.lm(max_iters=iters or o.factor_iters)
        self.update_map(ba, kfs)
        ba.close()
        return self.last_report

    def update_map(self, ba: LocalBA, kfs):
        """Mapper::UpdateMap (mapper.cpp:1141-1180): code, pose_wk, dpt_scale and UpdateDepth(...) -> dpt_map per keyframe."""
        poses, codes, scales, maps = ba.update_map()
        for n, kf in enumerate(kfs):
            kf.code = codes[n].copy()
            kf.pose_wk = poses[n]
            kf.dpt_scale = float(scales[n])
            kf.dpt_map_stored = maps[n].reshape(kf.video_mask.shape).copy()

    def close(self):
        for d in self._dev.values():
            d.close()
        self._dev = {}

"""Batched local bundle adjustment over the C ABI's problem object, plus the multi-GPU plumbing.

One process per GPU.  The distinct ordered pairs (i -> j), sorted by (i, j), are cut into equal runs, one per rank
(`shard_owners`, the rule of sage_ba_shard_plan): a keyframe's pairs are consecutive, so its maps are read by one GPU (two at a
cut), the loads differ by at most one pair, and a rank uploads only the keyframes its pairs touch (`needed_keyframes`).  Every rank linearises its factors
into ITS segment of the packed per-factor buffer [AtA | Atb | error | inliers]*, one in-place all-gather completes the buffer
everywhere, and every rank assembles (fixed order) and solves (block Cholesky over keyframes) the same normal equations --
the elimination is a 32-step dependency chain over the band of the covisibility graph, cheaper to repeat than to distribute
(DESIGN.md section 6).  The collective is issued by the library itself on the context's stream (NCCL bound at run time,
`LocalBA.enable_nccl`); torch.distributed only distributes the 128-byte communicator id.  A callback-based sum all-reduce
(`enable_allreduce`) remains for hosts without NCCL and for the gloo CPU tests.
"""
import ctypes as C

import numpy as np

from . import capi
from .ops import Context, DeviceKeyframe, SageError, _f, _p

F32 = np.float32


def shard_owners(pairs, world):
    """Owner rank of every ordered pair (i, j): the distinct pairs, sorted by (i, j), are cut into `world` equal runs
    (the rule of sage_ba_shard_plan).  Returns {(i, j): rank}."""
    uniq = sorted(set((int(i), int(j)) for i, j in pairs))
    P = len(uniq)
    return {pr: (0 if world <= 1 else idx * world // P) for idx, pr in enumerate(uniq)}


def shard_factors(factors, rank, world):
    """Indices of the factors (kind, i, j) owned by `rank`."""
    own = shard_owners([(i, j) for _, i, j in factors], world)
    return [f for f, (_, i, j) in enumerate(factors) if own[(i, j)] == rank]


def needed_keyframes(pairs, rank, world):
    """Keyframes whose device data `rank` needs: hosts and targets of the ordered pairs it owns (at least one keyframe, so that
    a rank without any pair still knows the shapes)."""
    own = shard_owners(pairs, world)
    need = set()
    for pr, r in own.items():
        if r == rank:
            need.update(pr)
    if not need and own:
        need.add(min(own)[0])
    return need


def factor_layout(kinds, C_code, owners=None, world=1):
    """Offsets of each factor's [AtA D*D | Atb D | error | inliers] block in the packed buffer: `world` equal segments,
    segment r holding rank r's factors in order of addition, segments padded to a multiple of 32 floats (world = 1: the
    order of addition).  kinds: sequence of 'photo' | 'geo' | 'reproj'.  Returns (offsets, dims, total)."""
    owners = [0] * len(kinds) if owners is None else list(owners)
    used = [0] * world
    local, dims = [], []
    for k, o in zip(kinds, owners):
        D = 14 + 2 * C_code if k == "geo" else 13 + C_code
        local.append(used[o])
        dims.append(D)
        used[o] += D * D + D + 2
    if world == 1:
        return local, dims, used[0]
    seg = (max(max(used), 4) + 31) // 32 * 32
    return [o * seg + l for o, l in zip(owners, local)], dims, seg * world


def pack_factor(buf, off, D, AtA, Atb, error, inliers):
    buf[off:off + D * D] = np.asarray(AtA, buf.dtype).reshape(-1)
    buf[off + D * D:off + D * D + D] = np.asarray(Atb, buf.dtype).reshape(-1)
    buf[off + D * D + D] = error
    buf[off + D * D + D + 1] = inliers


def variable_index(kind, i, j, c, K, C_code):
    """Global variable index of column c of a factor between keyframes i -> j.
    Global order: [pose_0..pose_{K-1} (6 each) | (code_k (C), scale_k) ...]."""
    cb_i, cb_j = 6 * K + i * (C_code + 1), 6 * K + j * (C_code + 1)
    if c < 6:
        return 6 * i + c
    if c < 12:
        return 6 * j + (c - 6)
    if kind == "geo":
        if c < 12 + C_code:
            return cb_i + (c - 12)
        if c < 12 + 2 * C_code:
            return cb_j + (c - 12 - C_code)
        return cb_i + C_code if c == 12 + 2 * C_code else cb_j + C_code
    if c < 12 + C_code:
        return cb_i + (c - 12)
    return cb_i + C_code


def assemble_dense(buf, factors, K, C_code, world=1):
    """Host restatement of the device assembly: dense fp64 (H, g, cost) from a packed factor buffer laid out for `world` ranks.
    factors: list of (kind, i, j).  Used by the gloo tests and as documentation of the layout."""
    n = K * (7 + C_code)
    H, g, cost = np.zeros((n, n)), np.zeros(n), 0.0
    own = shard_owners([(i, j) for _, i, j in factors], world)
    offs, dims, _ = factor_layout([f[0] for f in factors], C_code, [own[(f[1], f[2])] for f in factors], world)
    for (kind, i, j), off, D in zip(factors, offs, dims):
        idx = np.array([variable_index(kind, i, j, c, K, C_code) for c in range(D)])
        A = np.asarray(buf[off:off + D * D], np.float64).reshape(D, D)
        b = np.asarray(buf[off + D * D:off + D * D + D], np.float64)
        H[np.ix_(idx, idx)] += A
        np.add.at(g, idx, b)
        cost += float(buf[off + D * D + D])
    return H, g, cost


def allreduce_sum(tensor):
    """Sum-all-reduce through torch.distributed if a process group is up (NCCL or gloo); no-op otherwise."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


class _DevPtr:
    """Zero-copy torch view of a device buffer owned by libsage_ba (via __cuda_array_interface__)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class LocalBA:
    """Batched LM over K keyframes and a list of factors; Mapper::MappingStep's role (mapper.cpp:469-612)
    without ISAM2.  State = (pose_wk [K] as (R,t), code [K,C], dpt_scale [K])."""

    def __init__(self, ctx: Context, keyframes, rank=0, world=1, solver="auto"):
        """keyframes: DeviceKeyframe per keyframe; entries this rank's factors never touch may be None (see needed_keyframes)."""
        self.ctx = ctx
        self.kfs = list(keyframes)
        self.K = len(self.kfs)
        any_kf = next(k for k in self.kfs if k is not None)
        self.C = any_kf.C
        self.L = any_kf.L
        self._shape = (any_kf.H, any_kf.W)
        arr = (C.c_void_p * self.K)(*[(k.h if k is not None else None) for k in self.kfs])
        h = C.c_void_p()
        ctx.check(ctx.lib.sage_ba_problem_create(ctx.h, self.K, arr, C.byref(h)))
        self.h = h
        self.rank, self.world = rank, world
        ctx.check(ctx.lib.sage_ba_problem_set_shard(self.h, rank, world))
        ctx.check(ctx.lib.sage_ba_problem_set_solver(self.h, {"auto": 0, "block": 0, "schur": 1, "dense": 1, "banded": 2, "natural": 2}[solver]))
        self.factors = []
        self._cb = None
        self._comm = None
        self._views = {}

    # -- factor graph ------------------------------------------------------------------------------
    def add_photometric(self, i, j, weights):
        w = _f(weights)
        self.ctx.check(self.ctx.lib.sage_ba_problem_add_photometric(self.h, i, j, _p(w)))
        self.factors.append(("photo", i, j))

    def add_geometric(self, i, j, loss_param, weight):
        self.ctx.check(self.ctx.lib.sage_ba_problem_add_geometric(self.h, i, j, float(loss_param), float(weight)))
        self.factors.append(("geo", i, j))

    def add_reprojection(self, i, j, loc1d, homo, match2d, loss_param, weight):
        loc, hm, m2 = np.ascontiguousarray(loc1d, np.int32), _f(homo), _f(match2d)
        self.ctx.check(self.ctx.lib.sage_ba_problem_add_reprojection(self.h, i, j, _p(loc), _p(hm), _p(m2), len(loc),
                                                                     float(loss_param), float(weight)))
        self.factors.append(("reproj", i, j))

    def add_code_prior(self, k, weight, init_code=None):
        c = _f(init_code) if init_code is not None else np.zeros(self.C, F32)
        self.ctx.check(self.ctx.lib.sage_ba_problem_add_code_prior(self.h, k, _p(c), float(weight)))

    def add_scale_prior(self, k, init_scale, weight):
        self.ctx.check(self.ctx.lib.sage_ba_problem_add_scale_prior(self.h, k, float(init_scale), float(weight)))

    def fix(self, k, pose=True, scale=True):
        self.ctx.check(self.ctx.lib.sage_ba_problem_fix(self.h, k, int(pose), int(scale)))

    # -- state ---------------------------------------------------------------------------------------
    def set_state(self, poses, codes, scales, eps=1e-4):
        P = np.stack([np.concatenate([_f(R).reshape(-1), _f(t).reshape(-1)]) for R, t in poses]).astype(F32)
        c, s = _f(codes), _f(scales)
        self.ctx.check(self.ctx.lib.sage_ba_problem_set_state(self.h, _p(P), _p(c), _p(s), float(eps)))

    def get_state(self):
        P = np.zeros((self.K, 12), F32)
        c = np.zeros((self.K, self.C), F32)
        s = np.zeros((self.K,), F32)
        self.ctx.check(self.ctx.lib.sage_ba_problem_get_state(self.h, _p(P), _p(c), _p(s)))
        return [(P[k, :9].reshape(3, 3).copy(), P[k, 9:].copy()) for k in range(self.K)], c, s

    def update_map(self, want_depth=True):
        """Mapper::UpdateMap's payload (mapper.cpp:1141-1180): (poses, codes, scales, dpt_maps [K,HW]); the depth maps are
        UpdateDepth(...) evaluated on the device for the current estimate."""
        P = np.zeros((self.K, 12), F32)
        c = np.zeros((self.K, self.C), F32)
        s = np.zeros((self.K,), F32)
        HW = self._shape[0] * self._shape[1]
        d = np.zeros((self.K, HW), F32) if want_depth else None
        self.ctx.check(self.ctx.lib.sage_ba_problem_update_map(self.h, _p(P), _p(c), _p(s), _p(d), capi.HOST))
        return [(P[k, :9].reshape(3, 3).copy(), P[k, 9:].copy()) for k in range(self.K)], c, s, d

    @property
    def dim(self):
        return int(self.ctx.lib.sage_ba_problem_dim(self.h))

    @property
    def num_residuals(self):
        return int(self.ctx.lib.sage_ba_problem_num_residuals(self.h))

    # -- multi-GPU -----------------------------------------------------------------------------------
    def _buffer_view(self, which):
        import torch

        if which not in self._views:
            ptr, cnt = C.c_void_p(), C.c_size_t()
            fn = self.ctx.lib.sage_ba_problem_factor_buffer if which == "factor" else self.ctx.lib.sage_ba_problem_cost_buffer
            self.ctx.check(fn(self.h, C.byref(ptr), C.byref(cnt)))
            self._views[which] = (ptr.value, cnt.value,
                                  torch.as_tensor(_DevPtr(ptr.value, cnt.value), device=torch.device("cuda", self.ctx.device)))
        return self._views[which]

    def enable_nccl(self):
        """Create the library's own NCCL communicator (one per LocalBA) and hand it to the problem: from then on the exchange of
        the factor / cost buffers is issued from C++ on the context's stream.  torch.distributed only carries the 128-byte id."""
        import torch
        import torch.distributed as dist

        if self.world <= 1 or self._comm is not None:
            return
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            if self.ctx.lib.sage_ba_nccl_unique_id(buf) != 0:
                raise SageError("sage_ba_nccl_unique_id failed (libnccl not loadable)")
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        dev = torch.device("cuda", self.ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        ident = ident.to(dev)
        dist.broadcast(ident, 0)
        raw = bytes(ident.cpu().numpy().tobytes())
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sage_ba_comm_create(self.ctx.h, raw, self.rank, self.world, C.byref(h)))
        self._comm = h
        self.ctx.check(self.ctx.lib.sage_ba_problem_set_comm(self.h, h))

    def exchange(self, which="factor"):
        """Complete the factor / cost buffer on every rank (library collective or callback; no-op for world = 1)."""
        self.ctx.check(self.ctx.lib.sage_ba_problem_exchange(self.h, 0 if which == "factor" else 1))

    def factor_offsets(self):
        """(offsets, cost_offsets, owners) per factor in order of addition, as laid out by the library."""
        n = len(self.factors)
        a, b, c = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        self.ctx.check(self.ctx.lib.sage_ba_problem_factor_offsets(self.h, a, b, c))
        return list(a), list(b), list(c)

    def solver_info(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self.ctx.check(self.ctx.lib.sage_ba_problem_solver_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"blocks": a.value, "fill_blocks": b.value, "depth": c.value}

    def deterministic(self, on=True):
        """CTA decomposition independent of the rank count: bit-identical LM trajectories for every number of GPUs."""
        self.ctx.check(self.ctx.lib.sage_ba_problem_set_deterministic(self.h, int(on)))

    def relinearize_always(self, always=True):
        self.ctx.check(self.ctx.lib.sage_ba_problem_set_relinearize_always(self.h, int(always)))

    def _ensure_collective(self):
        if self.world > 1 and self._comm is None and self._cb is None:
            import torch.distributed as dist

            if dist.is_initialized() and dist.get_backend() == "nccl":
                self.enable_nccl()
            else:
                self.enable_allreduce()

    def enable_allreduce(self):
        """Install the all-reduce callback the C++ LM loop calls once per linearisation and once per trial."""
        views = {self._buffer_view(w)[0]: self._buffer_view(w)[2] for w in ("factor", "cost")}

        import torch

        ext = torch.cuda.ExternalStream(self.ctx.stream_handle, device=torch.device("cuda", self.ctx.device))

        def cb(ptr, count, user):
            try:
                with torch.cuda.stream(ext):  # the collective must be ordered with the library's kernels on the context's stream
                    allreduce_sum(views[ptr][:count])
                return 0
            except Exception as e:  # pragma: no cover
                print("all-reduce callback failed:", e)
                return 1

        self._cb = capi.ALLREDUCE_FN(cb)
        self.ctx.check(self.ctx.lib.sage_ba_problem_set_allreduce(self.h, self._cb, None))

    # -- stages (sage_ba_problem_*) --------------------------------------------------------------------
    def linearize(self, reduce=True):
        self.ctx.check(self.ctx.lib.sage_ba_problem_linearize(self.h))
        if reduce and self.world > 1:
            self._ensure_collective()
            self.exchange("factor")

    def assemble(self, want_matrix=False):
        cost = C.c_double(0)
        if want_matrix:
            n = self.dim
            H, g = np.zeros((n, n)), np.zeros(n)
            self.ctx.check(self.ctx.lib.sage_ba_problem_assemble(self.h, _p(H), _p(g), C.byref(cost)))
            return H, g, cost.value
        self.ctx.check(self.ctx.lib.sage_ba_problem_assemble(self.h, None, None, C.byref(cost)))
        return cost.value

    def solve(self, damp, want_delta=False):
        if want_delta:
            d = np.zeros(self.dim)
            self.ctx.check(self.ctx.lib.sage_ba_problem_solve(self.h, float(damp), _p(d)))
            return d
        self.ctx.check(self.ctx.lib.sage_ba_problem_solve(self.h, float(damp), None))

    def evaluate(self, candidate=True, reduce=True):
        self.ctx.check(self.ctx.lib.sage_ba_problem_evaluate(self.h, int(candidate)))
        if reduce and self.world > 1:
            self._ensure_collective()
            self.exchange("cost")
        cost = C.c_double(0)
        self.ctx.check(self.ctx.lib.sage_ba_problem_cost(self.h, int(candidate), C.byref(cost)))
        return cost.value

    def accept(self):
        self.ctx.check(self.ctx.lib.sage_ba_problem_accept(self.h))

    def factor_buffer(self):
        """Host copy of the packed per-factor outputs (after linearize)."""
        self.ctx.synchronize()
        return self._buffer_view("factor")[2].cpu().numpy()

    def lm(self, max_iters=10, init_damp=1e-4, min_damp=1e-6, max_damp=1e2, damp_dec_factor=10.0, damp_inc_factor=10.0,
           min_rel_decrease=1e-6, max_trials=8):
        opt = capi.LMOptions(max_iters, init_damp, min_damp, max_damp, damp_dec_factor, damp_inc_factor, min_rel_decrease,
                             max_trials)
        rep = capi.LMReport()
        self._ensure_collective()
        self.ctx.check(self.ctx.lib.sage_ba_problem_lm(self.h, C.byref(opt), C.byref(rep)))
        return {k: getattr(rep, k) for k, _ in capi.LMReport._fields_}

    def lm_step(self, damp, min_damp=1e-6, max_damp=1e2, damp_dec_factor=10.0, damp_inc_factor=10.0):
        """One LM iteration (linearise, assemble, solve, evaluate the candidate, accept / reject) with a single host
        synchronisation.  Returns (cost, candidate_cost, accepted, new_damp)."""
        d, c0, c1, acc = C.c_double(damp), C.c_double(0), C.c_double(0), C.c_int(0)
        self._ensure_collective()
        self.ctx.check(self.ctx.lib.sage_ba_problem_lm_step(self.h, C.byref(d), min_damp, max_damp, damp_dec_factor, damp_inc_factor,
                                                            C.byref(c0), C.byref(c1), C.byref(acc)))
        return c0.value, c1.value, bool(acc.value), d.value

    def profile(self, enable=True):
        self.ctx.check(self.ctx.lib.sage_ba_problem_profile(self.h, int(enable)))

    def profile_read(self, reset=True):
        """{kind: (total_ms, launches)} measured with CUDA events on the context's stream."""
        n = len(capi.PROF_KINDS)
        ms, cnt = (C.c_double * n)(), (C.c_long * n)()
        self.ctx.check(self.ctx.lib.sage_ba_problem_profile_read(self.h, ms, cnt, int(reset)))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(capi.PROF_KINDS)}

    def shard_counts(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self.ctx.lib.sage_ba_problem_shard_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"photo": a.value, "geo": b.value, "reproj": c.value}

    def close(self):
        if self.h:
            self.ctx.lib.sage_ba_problem_destroy(self.h)
            self.h = None
        if self._comm is not None:
            self.ctx.lib.sage_ba_comm_destroy(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Host mirror of the reference's GTSAM factor classes for the hot path:

    PhotometricFactor   core/gtsam/photometric_factor.cpp:72 (error), :106-219 (linearize), :264-311
    GeometricFactor     core/gtsam/geometric_factor.cpp:41, :67-233, :300-356
    ReprojectionFactor  core/gtsam/reprojection_factor.cpp:201, :227-317, :332-396
    MatchGeometryFactor core/gtsam/match_geometry_factor.cpp (same structure; 3-D point-to-point term of loop closure / scale init)
    cycle_matches       the matching part of both constructors (reprojection_factor.cpp:36-112, match_geometry_factor.cpp:40-118)

Same responsibilities as the reference classes: unpack `Values` (pose_wk as (R, t), code, scale), build the
relative pose T10 = T1^-1 T0 in fp32 (photometric_factor.cpp:280-281), call the operator (here: the C ABI),
cast AtA/Atb to fp64, apply NearestPsd (core/mapping/mapping_utils.h:104-128) and slice the upper-triangular
block list of a gtsam::HessianFactor(keys, Gs, gs, f = error)  [E(x) = 1/2 x^T G x - x^T g + 1/2 f].

NearestPsd in the reference forms V^T S V instead of V S V^T, which is not a projection.
`psd="reference"` reproduces that bit of behaviour, `psd="exact"` is Higham's projection, `psd="none"` skips it.
GTSAM itself is not available here, so HessianFactor is a plain record with the same fields.
"""
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import ops
from .frames import std_shuffle

F32 = np.float32


def pose_key(k):
    """gtsam::Symbol('p', id)  (core/gtsam/gtsam_utils.h:10-13)"""
    return ("p", int(k))


def code_key(k):
    return ("c", int(k))


def scale_key(k):
    return ("s", int(k))


class Values(dict):
    """gtsam::Values stand-in: {pose_key: (R [3,3], t [3]), code_key: code [C] (double), scale_key: float}."""


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


def nearest_psd(M, mode="reference"):
    """NearestPsd (core/mapping/mapping_utils.h:104-128) in fp64.  mode "reference": the reference's numbers, V^T S V quirk and
    Eigen's sign conventions included (eigen_nearest_psd, pinned by tests/golden/host_pins.npz); "exact": Higham's projection
    V S V^T; "none": pass-through."""
    M = np.asarray(M, dtype=np.float64)
    if mode == "none":
        return M
    if mode == "reference":
        return eigen_nearest_psd(M)
    B = (M + M.T) / 2
    _, s, Vt = np.linalg.svd(B)
    V = Vt.T
    A2 = (B + V @ np.diag(s) @ V.T) / 2
    A3 = (A2 + A2.T) / 2
    k, I = 1, np.eye(M.shape[0])
    while True:
        try:
            np.linalg.cholesky(A3)
            return A3
        except np.linalg.LinAlgError:
            A3 = A3 + I * (-np.linalg.eigvalsh(A3).min() * k + 1e-15)
            k *= 2


@dataclass
class HessianFactor:
    """gtsam::HessianFactor(keys, Gs, gs, f): Gs is the upper-triangular block list in row-major block order."""
    keys: List[Tuple[str, int]]
    dims: List[int]
    Gs: List[np.ndarray]
    gs: List[np.ndarray]
    f: float

    def information(self):
        """Dense symmetric G and g in the factor's key order."""
        off = np.concatenate([[0], np.cumsum(self.dims)])
        n = int(off[-1])
        G, g = np.zeros((n, n)), np.zeros(n)
        it = iter(self.Gs)
        for a in range(len(self.dims)):
            for b in range(a, len(self.dims)):
                blk = next(it)
                G[off[a]:off[a + 1], off[b]:off[b + 1]] = blk
                G[off[b]:off[b + 1], off[a]:off[a + 1]] = blk.T
            g[off[a]:off[a + 1]] = self.gs[a].reshape(-1)
        return G, g


def _partition(AtA, Atb, keys, dims, error, psd):
    G = nearest_psd(np.asarray(AtA, np.float64), psd)
    g = np.asarray(Atb, np.float64).reshape(-1)
    off = np.concatenate([[0], np.cumsum(dims)])
    Gs, gs = [], []
    for a in range(len(dims)):
        for b in range(a, len(dims)):
            Gs.append(G[off[a]:off[a + 1], off[b]:off[b + 1]].copy())
        gs.append(g[off[a]:off[a + 1]].copy())
    return HessianFactor(list(keys), list(dims), Gs, gs, float(error))


def _rel(p0, p1):
    R0, t0 = np.asarray(p0[0], F32), np.asarray(p0[1], F32)
    R1, t1 = np.asarray(p1[0], F32), np.asarray(p1[1], F32)
    return (R1.T @ R0).astype(F32), (R1.T @ (t0 - t1)).astype(F32), R0, t0, R1, t1


class PhotometricFactor:
    """keys (pose0, pose1, code0, scale0); 10 blocks (photometric_factor.cpp:151-218)."""

    def __init__(self, ctx, kf, fr, factor_weights, dpt_eps=1e-4, psd="reference"):
        self.ctx, self.kf, self.fr = ctx, kf, fr  # DeviceKeyframe handles; kf.kf.id / fr.kf.id are the frame ids
        self.weights, self.eps, self.psd = np.asarray(factor_weights, F32), dpt_eps, psd
        self.keys = [pose_key(kf.kf.id), pose_key(fr.kf.id), code_key(kf.kf.id), scale_key(kf.kf.id)]
        self.error_ = None

    def dim(self):
        return 13 + self.kf.C

    def _unpack(self, c):
        return c[self.keys[0]], c[self.keys[1]], np.asarray(c[self.keys[2]], F32), float(c[self.keys[3]])

    def error(self, c):
        p0, p1, code, s = self._unpack(c)
        R10, t10, *_ = _rel(p0, p1)
        e, _ = ops.photometric_error_calculate(self.ctx, self.kf, self.fr, R10, t10, code, s, self.eps, self.weights)
        return float(e)

    def linearize(self, c):
        p0, p1, code, s = self._unpack(c)
        R10, t10, R0, t0, R1, t1 = _rel(p0, p1)
        AtA, Atb, e, _ = ops.photometric_jac_error_calculate(self.ctx, self.kf, self.fr, R10, t10, R0, t0, R1, t1, code, s,
                                                             self.eps, self.weights)
        self.error_ = e
        return _partition(AtA, Atb, self.keys, [6, 6, self.kf.C, 1], e, self.psd)


class GeometricFactor:
    """keys (pose0, pose1, code0, code1, scale0, scale1); 21 blocks (geometric_factor.cpp:122-233)."""

    def __init__(self, ctx, kf0, kf1, factor_weight, loss_param, dpt_eps=1e-4, psd="reference"):
        self.ctx, self.kf0, self.kf1 = ctx, kf0, kf1
        self.weight, self.loss_param, self.eps, self.psd = factor_weight, loss_param, dpt_eps, psd
        i, j = kf0.kf.id, kf1.kf.id
        self.keys = [pose_key(i), pose_key(j), code_key(i), code_key(j), scale_key(i), scale_key(j)]

    def dim(self):
        return 14 + 2 * self.kf0.C

    def _unpack(self, c):
        k = self.keys
        return c[k[0]], c[k[1]], np.asarray(c[k[2]], F32), np.asarray(c[k[3]], F32), float(c[k[4]]), float(c[k[5]])

    def error(self, c):
        p0, p1, c0, c1, s0, s1 = self._unpack(c)
        R10, t10, *_ = _rel(p0, p1)
        e, _ = ops.geometric_error_calculate(self.ctx, self.kf0, self.kf1, R10, t10, c0, c1, s0, s1, self.eps, self.loss_param,
                                             self.weight)
        return float(e)

    def linearize(self, c):
        p0, p1, c0, c1, s0, s1 = self._unpack(c)
        R10, t10, R0, t0, R1, t1 = _rel(p0, p1)
        AtA, Atb, e, _ = ops.geometric_jac_error_calculate(self.ctx, self.kf0, self.kf1, R10, t10, R0, t0, R1, t1, c0, c1, s0, s1,
                                                           self.eps, self.loss_param, self.weight)
        C = self.kf0.C
        return _partition(AtA, Atb, self.keys, [6, 6, C, C, 1, 1], e, self.psd)


class ReprojectionFactor:
    """keys (pose0, pose1, code0, scale0); matches are found once by the constructor in the reference
    (reprojection_factor.cpp:7-193, section 8 f3 = next) and are handed in here."""

    def __init__(self, ctx, kf, fr, matched_locations_1d_0, matched_locations_homo_0, matched_locations_2d_1, factor_weight,
                 loss_param, inlier_multiplier=1.0, dpt_eps=1e-4, psd="reference"):
        self.ctx, self.kf, self.fr = ctx, kf, fr
        self.loc, self.homo, self.m2d = matched_locations_1d_0, matched_locations_homo_0, matched_locations_2d_1
        self.weight, self.loss_param, self.eps, self.psd = inlier_multiplier * factor_weight, loss_param, dpt_eps, psd
        self.keys = [pose_key(kf.kf.id), pose_key(fr.kf.id), code_key(kf.kf.id), scale_key(kf.kf.id)]

    def dim(self):
        return 13 + self.kf.C

    def error(self, c):
        p0, p1 = c[self.keys[0]], c[self.keys[1]]
        R10, t10, *_ = _rel(p0, p1)
        e, _ = ops.reprojection_error_calculate(self.ctx, self.kf, R10, t10, np.asarray(c[self.keys[2]], F32), float(c[self.keys[3]]),
                                                self.loc, self.homo, self.m2d, self.eps, self.loss_param, self.weight)
        return float(e)

    def linearize(self, c):
        p0, p1 = c[self.keys[0]], c[self.keys[1]]
        R10, t10, R0, t0, R1, t1 = _rel(p0, p1)
        AtA, Atb, e, _ = ops.reprojection_jac_error_calculate(self.ctx, self.kf, R10, t10, R0, t0, R1, t1,
                                                              np.asarray(c[self.keys[2]], F32), float(c[self.keys[3]]), self.loc,
                                                              self.homo, self.m2d, self.eps, self.loss_param, self.weight)
        return _partition(AtA, Atb, self.keys, [6, 6, self.kf.C, 1], e, self.psd)


def cycle_matches(ctx, kf_rec, fr_rec, num_keypoints, cyc_consis_thresh, seed=None, libstdcxx="13"):
    """The descriptor-matching part of the ReprojectionFactor / MatchGeometryFactor constructors
    (reprojection_factor.cpp:36-112): draw `num_keypoints` of kf's valid locations, cycle-match their descriptors against fr
    (ops.cycle_feature_matching = row f3 on the device), keep the cycle-consistent ones.

    kf_rec / fr_rec are frames.Keyframe records with feat_desc.  Returns None without inliers, else a dict with the
    reference's member names: matched_locations_1d_0 [M] int32, matched_locations_homo_0 [M,3], matched_locations_1d_1 [M] int32,
    matched_locations_2d_1 [M,2], matched_locations_homo_1 [M,3], desc_inlier_ratio.
    The keypoint draw is std::shuffle(iota(n), std::mt19937(kf.id * fr.id)) like the reference (frames.std_shuffle; `libstdcxx`
    selects the library generation, see there).  Not reproduced: the TEASER++ filtering that follows
    (reprojection_factor.cpp:136-186, third-party)."""
    n = len(kf_rec.sampled_locations_1d)
    K = min(int(num_keypoints), n)
    idx = std_shuffle(n, kf_rec.id * fr_rec.id if seed is None else seed, libstdcxx)[:K]
    kp = np.asarray(kf_rec.sampled_locations_1d)[idx]
    r = ops.cycle_feature_matching(ctx, kf_rec.feat_desc, fr_rec.feat_desc, kp, cyc_consis_thresh)
    sel = r["inlier_within_keypoint_indexes"]
    if len(sel) == 0:
        return None
    cam = fr_rec.camera_pyramid[0]
    uv = r["matched_locations_2d_1"].astype(F32)
    homo1 = np.stack([(uv[:, 0] - F32(cam[2])) / F32(cam[0]), (uv[:, 1] - F32(cam[3])) / F32(cam[1]), np.ones(len(uv), F32)], 1)
    return {"matched_locations_1d_0": kp[sel].astype(np.int32),
            "matched_locations_homo_0": np.asarray(kf_rec.sampled_locations_homo)[idx[sel]].astype(F32),
            "matched_locations_1d_1": r["matched_locations_1d_1"].astype(np.int32), "matched_locations_2d_1": uv,
            "matched_locations_homo_1": homo1.astype(F32), "desc_inlier_ratio": len(sel) / float(K)}


class MatchGeometryFactor:
    """keys (pose0, pose1, code0, code1, scale0, scale1); 3-D point-to-point term between matched keypoints with the depths of
    both keyframes (core/gtsam/match_geometry_factor.cpp; kernels cuda/match_geometry_factor_kernels.cpp:1567-1823)."""

    def __init__(self, ctx, kf0, kf1, matches, factor_weight, loss_param, robust_loss_type="fair", inlier_multiplier=1.0, psd="reference"):
        self.ctx, self.kf0, self.kf1, self.m = ctx, kf0, kf1, matches
        self.weight, self.loss_param, self.loss_type, self.psd = inlier_multiplier * factor_weight, loss_param, robust_loss_type, psd
        i, j = kf0.kf.id, kf1.kf.id
        self.keys = [pose_key(i), pose_key(j), code_key(i), code_key(j), scale_key(i), scale_key(j)]

    def dim(self):
        return 14 + 2 * self.kf0.C

    def _args(self, c):
        m = self.m
        return (np.asarray(c[self.keys[2]], F32), np.asarray(c[self.keys[3]], F32), m["matched_locations_homo_0"],
                m["matched_locations_homo_1"], m["matched_locations_1d_0"], m["matched_locations_1d_1"], float(c[self.keys[4]]),
                float(c[self.keys[5]]), self.loss_param, self.weight, self.loss_type)

    def error(self, c):
        R10, t10, *_ = _rel(c[self.keys[0]], c[self.keys[1]])
        return float(ops.match_geometry_error_calculate(self.ctx, self.kf0, self.kf1, R10, t10, *self._args(c)))

    def linearize(self, c):
        R10, t10, R0, t0, R1, t1 = _rel(c[self.keys[0]], c[self.keys[1]])
        AtA, Atb, e = ops.match_geometry_jac_error_calculate(self.ctx, self.kf0, self.kf1, R10, t10, R0, t0, R1, t1, *self._args(c))
        C = self.kf0.C
        return _partition(AtA, Atb, self.keys, [6, 6, C, C, 1, 1], e, self.psd)

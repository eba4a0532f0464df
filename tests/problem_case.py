"""A small multi-keyframe BA problem shared by the GPU solver test and the gloo test: factor list + the oracle's
per-factor outputs packed like the device factor buffer."""
import numpy as np

import helpers
import oracle as O
import sage_slam_b200 as sage
from sage_slam_b200 import local_ba

PRM = dict(W=64, H=48, L=3, F=16, C=8, num_samples=None, mask="full", seed=21)
CODE_W, SCALE_W = 1e-3, 1e-2


def build(num_kf=4):
    kfs = sage.synthetic.make_scene(num_kf=num_kf, **PRM)
    rng = np.random.default_rng(77)
    for k in kfs:
        k.code = (0.1 * rng.standard_normal(PRM["C"])).astype(np.float32)
        k.dpt_scale = float(np.float32(1.0 + 0.05 * rng.standard_normal()))
    pairs = sage.synthetic.ordered_pairs(kfs)
    factors = [("photo", i, j) for i, j in pairs] + [("geo", i, j) for i, j in pairs] + [("reproj", i, j) for i, j in pairs]
    return kfs, pairs, factors


def geo_loss(kfs):
    return float(0.03 * np.mean(kfs[0].dpt_map_bias.astype(np.float64) ** 2))


def matches(kfs, i, j):
    return sage.synthetic.make_matches(kfs[i], kfs[j], M=64)


def oracle_factor(kfs, factor, dtype=np.float32):
    kind, i, j = factor
    a = helpers.case_args(kfs, i, j)
    if kind == "photo":
        return O.photometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                       a["mask1"], a["loc1d"], a["homo"], a["feat0"], a["feat1"], a["grad1"], a["level_offsets"],
                                       a["scale0"], a["cams"], a["eps"], a["weights"], dtype=dtype)
    if kind == "geo":
        return O.geometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                     a["dpt1"], a["dgrad1"], a["basis1"], a["mask1"], a["loc1d"], a["homo"], a["scale0"],
                                     a["scale1"], a["cam"], a["eps"], geo_loss(kfs), 0.1, dtype=dtype)
    loc, homo, uv = matches(kfs, i, j)
    return O.reprojection_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"], loc,
                                    homo, uv, a["scale0"], a["cam"], a["eps"], 0.03 * PRM["W"] ** 2, 0.1, dtype=dtype)


def oracle_buffer(kfs, factors, owned=None, world=1):
    """Packed fp32 factor buffer (laid out for `world` ranks) with only the `owned` factor indices filled (others zero), like
    one rank's shard before the exchange."""
    own = local_ba.shard_owners([(i, j) for _, i, j in factors], world)
    owners = [own[(f[1], f[2])] for f in factors]
    offs, dims, total = local_ba.factor_layout([f[0] for f in factors], PRM["C"], owners, world)
    buf = np.zeros(total, np.float32)
    for f, (fac, off, D) in enumerate(zip(factors, offs, dims)):
        if owned is not None and f not in owned:
            continue
        A, b, e, n = oracle_factor(kfs, fac)
        local_ba.pack_factor(buf, off, D, A, b, e, n)
    return buf


def add_priors_dense(H, g, kfs):
    """CodeFactor / ScaleFactor contributions (gtsam/code_factor.cpp:42-104, scale_factor.cpp:115-130), fp64."""
    K, C = len(kfs), PRM["C"]
    cost = 0.0
    for k, kf in enumerate(kfs):
        cb = 6 * K + k * (C + 1)
        diff = -kf.code.astype(np.float64)
        H[cb:cb + C, cb:cb + C] += CODE_W * np.eye(C)
        g[cb:cb + C] += CODE_W * diff
        cost += CODE_W * float(np.mean(diff ** 2))
        s = float(kf.dpt_scale)
        d = np.log(1.0) - np.log(s)
        H[cb + C, cb + C] += SCALE_W / (s * s)
        g[cb + C] += SCALE_W / s * d
        cost += SCALE_W * d * d
    return cost


def solve_dense(H, g, damp, fixed):
    Hd = H + damp * np.diag(np.diag(H))
    gd = g.copy()
    for v in fixed:
        Hd[v, :] = 0
        Hd[:, v] = 0
        Hd[v, v] = 1
        gd[v] = 0
    return np.linalg.solve(Hd, gd)

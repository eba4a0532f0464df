"""CPU: the oracle's Jacobians against numerical derivatives of the oracle's own error -- the check the reference's authors
left commented out (core/gtsam/photometric_factor.cpp:124-143).  With error(x + d) ~ error(x) + d^T AtA d - 2 Atb^T d the
gradient of the error is -2 Atb, for the robust (Cauchy / Fair) factors too (their rows carry the IRLS weight).  This pins
the restated Jacobian algebra independently of the reference goldens; fp64 instantiation of the oracle, central differences."""
import copy

import numpy as np
import pytest

import helpers
import oracle as O

EPS_FD = 2e-4


def _args64(kfs):
    """case_args with the poses (and everything derived from them) in fp64."""
    a = helpers.case_args(kfs)
    R0, t0 = (np.asarray(x, np.float64) for x in kfs[0].pose_wk)
    R1, t1 = (np.asarray(x, np.float64) for x in kfs[1].pose_wk)
    a.update(R0=R0, t0=t0, R1=R1, t1=t1, R10=R1.T @ R0, t10=R1.T @ (t0 - t1))
    return a


def _perturbed(kfs, which, k, h, C):
    """State vector layout of the reference factors: pose0 (v, w) | pose1 | code0 | [code1] | scale0 | [scale1]."""
    kf = copy.deepcopy(kfs)
    d = np.zeros(6)
    if which == "pose0" or which == "pose1":
        i = 0 if which == "pose0" else 1
        d[k] = h
        R, t = O.retract(np.asarray(kf[i].pose_wk[0], np.float64), np.asarray(kf[i].pose_wk[1], np.float64), d)
        kf[i].pose_wk = (R, t)
    elif which in ("code0", "code1"):
        i = 0 if which == "code0" else 1
        c = kf[i].code.astype(np.float64)
        c[k] += h
        kf[i].code = c
    else:
        i = 0 if which == "scale0" else 1
        kf[i].dpt_scale = float(kf[i].dpt_scale) + h
    return kf


def _check(kfs, layout, jac_fn, err_fn, tol):
    A, b, e = jac_fn(_args64(kfs))[:3]
    b = np.asarray(b, np.float64).reshape(-1)
    C = kfs[0].dpt_jac_code.shape[1]
    g_fd = []
    for which, n in layout:
        for k in range(n):
            ep = err_fn(_args64(_perturbed(kfs, which, k, +EPS_FD, C)))
            em = err_fn(_args64(_perturbed(kfs, which, k, -EPS_FD, C)))
            g_fd.append((ep - em) / (2 * EPS_FD))
    g_fd = np.array(g_fd)
    assert len(g_fd) == len(b)
    scale = np.abs(b).max()
    assert np.abs(g_fd + 2 * b).max() <= tol * 2 * scale, (np.abs(g_fd + 2 * b).max(), scale)
    # the Gauss-Newton matrix is symmetric positive semi-definite
    A = np.asarray(A, np.float64)
    assert np.abs(A - A.T).max() <= 1e-9 * np.abs(A).max() and np.linalg.eigvalsh((A + A.T) / 2).min() >= -1e-9 * np.abs(A).max()


@pytest.fixture(scope="module")
def kfs():
    k = helpers.build_case("small_c8_f16")
    # a state away from zero so that every block of the Jacobian matters
    rng = np.random.default_rng(4)
    for x in k:
        x.code = (0.05 * rng.standard_normal(x.code.shape)).astype(np.float32)
    k[0].dpt_scale, k[1].dpt_scale = 1.03, 0.98
    # keep only KF0 samples from the central half of the image: every warped sample then stays valid under the perturbations,
    # so the normalised masked sums are differentiable (a sample crossing the image border changes them by a jump, which a
    # finite difference reads as a spurious boundary-flux gradient of the same size as the true one)
    H, W = k[0].video_mask.shape
    loc = k[0].sampled_locations_1d
    y, x = loc // W, loc % W
    keep = (x > W // 4) & (x < 3 * W // 4) & (y > H // 4) & (y < 3 * H // 4)
    k[0].sampled_locations_1d = loc[keep]
    k[0].sampled_locations_homo = k[0].sampled_locations_homo[keep]
    return k


def test_photometric_jacobian_is_the_gradient_of_the_error(kfs):
    C = kfs[0].dpt_jac_code.shape[1]

    def jac(a):
        return O.photometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"], a["mask1"],
                                       a["loc1d"], a["homo"], a["feat0"], a["feat1"], a["grad1"], a["level_offsets"], a["scale0"],
                                       a["cams"], a["eps"], a["weights"], dtype=np.float64)

    def err(a):
        return float(O.photometric_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], a["mask1"], a["loc1d"], a["homo"], a["feat0"],
                                         a["feat1"], a["level_offsets"], a["scale0"], a["cams"], a["eps"], a["weights"], dtype=np.float64)[0])

    # the reference differentiates through a bilinearly sampled CENTRAL-DIFFERENCE gradient pyramid, which underestimates the
    # slope of the bilinear feature interpolant by the sinc factor of the feature wavelength (7-8 % on these textures):
    # agreement to ~10 % of the largest component, with the right sign on every component, is what the model allows
    _check(kfs, [("pose0", 6), ("pose1", 6), ("code0", C), ("scale0", 1)], jac, err, tol=0.12)


def test_geometric_jacobian_is_the_gradient_of_the_error(kfs):
    C = kfs[0].dpt_jac_code.shape[1]

    def jac(a):
        return O.geometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"], a["dpt1"],
                                     a["dgrad1"], a["basis1"], a["mask1"], a["loc1d"], a["homo"], a["scale0"], a["scale1"], a["cam"],
                                     a["eps"], a["geo_loss"], a["geo_weight"], dtype=np.float64)

    def err(a):
        return float(O.geometric_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], a["dpt1"], a["mask1"], a["loc1d"], a["homo"],
                                       a["scale0"], a["cam"], a["eps"], a["geo_loss"], a["geo_weight"], dtype=np.float64)[0])

    _check(kfs, [("pose0", 6), ("pose1", 6), ("code0", C), ("code1", C), ("scale0", 1), ("scale1", 1)], jac, err, tol=0.08)


def test_reprojection_jacobian_is_the_gradient_of_the_error(kfs):
    C = kfs[0].dpt_jac_code.shape[1]
    ma = helpers.match_args(kfs)

    def jac(a):
        return O.reprojection_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"], ma["mloc"],
                                        ma["mhomo"], ma["m2d"], a["scale0"], a["cam"], a["eps"], a["rep_loss"], a["rep_weight"],
                                        dtype=np.float64)

    def err(a):
        return float(O.reprojection_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], ma["mloc"], ma["mhomo"], ma["m2d"],
                                          a["scale0"], a["cam"], a["eps"], a["rep_loss"], a["rep_weight"], dtype=np.float64)[0])

    _check(kfs, [("pose0", 6), ("pose1", 6), ("code0", C), ("scale0", 1)], jac, err, tol=1e-3)

"""Which lineariser is closer to the fp64 truth as the number of sequential accumulations grows?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import helpers, oracle as O
import sage_slam_b200 as sage
from sage_slam_b200 import capi, ops
lib = capi.load()
ctx = sage.Context(0)
for (W, H, ns) in [(160, 128, 3000), (160, 128, None), (320, 256, None)]:
    prm = dict(W=W, H=H, L=3, F=16, C=32, seed=41)
    kfs = sage.synthetic.make_scene(num_kf=2, mask="full", num_samples=ns, **prm)
    rng = np.random.default_rng(6)
    for k in kfs:
        k.code = (0.2 * rng.standard_normal(32)).astype(np.float32)
        k.dpt_scale = float(np.float32(1.0 + 0.05 * rng.standard_normal()))
    a = helpers.case_args(kfs)
    Ao, bo, eo, no = O.geometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                           a["dpt1"], a["dgrad1"], a["basis1"], a["mask1"], a["loc1d"], a["homo"], a["scale0"], a["scale1"], a["cams"][0] if "cams" in a else a["cam"],
                                           a["eps"], a["geo_loss"], a["geo_weight"], dtype=np.float64)
    d0, d1 = sage.DeviceKeyframe(ctx, kfs[0]), sage.DeviceKeyframe(ctx, kfs[1])
    Ap, bp, ep, np_ = O.photometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                             a["mask1"], a["loc1d"], a["homo"], a["feat0"], a["feat1"], a["grad1"], a["level_offsets"],
                                             a["scale0"], a["cams"], a["eps"], a["weights"], dtype=np.float64)
    for sl in (None, "1", "4"):
        if sl: os.environ["SAGE_BA_SLICES"] = sl
        else: os.environ.pop("SAGE_BA_SLICES", None)
        A, b, e, n = ops.photometric_jac_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"], a["scale0"], a["eps"], a["weights"])
        ea, eb = helpers.block_errors("photo", A, b, Ap, bp, 32)
        dg = np.diag(A).astype(np.float64) / np.diag(Ap) - 1
        print(f"{W}x{H} N={len(a['loc1d'])} PHOTO slices={sl}: worst block err vs fp64 oracle {max(list(ea.values())+list(eb.values())):.2e}  diag rel bias mean {dg[np.isfinite(dg)].mean():+.2e}")
    for on in (0, 1):
        for sl in (None, "1", "4"):
            lib.sage_ba_set_geometric_tcgen05(on)
            if sl: os.environ["SAGE_BA_SLICES"] = sl
            else: os.environ.pop("SAGE_BA_SLICES", None)
            A, b, e, n = ops.geometric_jac_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"], a["code1"],
                                                           a["scale0"], a["scale1"], a["eps"], a["geo_loss"], a["geo_weight"])
            ea, eb = helpers.block_errors("geo", A, b, Ao, bo, 32)
            dg = np.diag(A).astype(np.float64) / np.diag(Ao) - 1
            print(f"{W}x{H} N={len(a['loc1d'])} tc={on} slices={sl}: worst block err vs fp64 oracle {max(list(ea.values())+list(eb.values())):.2e}  diag rel bias mean {dg[np.isfinite(dg)].mean():+.2e}  err {abs(e-eo)/eo:.1e} n {n} {no}")

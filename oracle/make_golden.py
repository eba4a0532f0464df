"""Generate tests/golden/*.npz by running the REFERENCE's own CUDA kernels (oracle/_ref, compiled
unmodified from /root/reference by build_ref.py) on the seeded cases of tests/helpers.py.

Must run on a GPU box:   gpurun -- 'python oracle/make_golden.py'   -> gpurun_out/golden/*.npz, which
are then copied into tests/golden/ and committed.  Only outputs are stored (inputs are regenerated
from the seeds); `sig` pins the inputs by checksum.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402
import helpers  # noqa: E402


def T(a, dev, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dtype)


def run_case(name, far, mod=None):
    dev = torch.device("cuda:0")
    assert torch.backends.cuda.matmul.allow_tf32 is False
    kfs = helpers.build_case(name, far=far)
    a = helpers.case_args(kfs)
    ta = helpers.tracker_args(kfs, a)
    ma = helpers.match_args(kfs)
    mod = mod or build_ref.load(a["C"], a["F"])
    cam = [float(x) for x in a["cam"]]
    L = a["L"]
    # the depth basis as the reference hands it over: [HW, C] view with strides (1, HW)
    jac0 = T(np.ascontiguousarray(a["jac0"].T), dev).t()
    assert jac0.stride() == (1, a["H"] * a["W"])
    R = {k: T(a[k], dev) for k in ("R10", "t10", "R0", "t0", "R1", "t1", "bias0", "code0", "mask1", "homo", "feat0", "feat1",
                                   "grad1", "dpt1", "dgrad1", "basis1")}
    loc64 = T(a["loc1d"], dev, torch.int64)
    loc32 = loc64.to(torch.int32)
    lo = T(a["level_offsets"], dev, torch.int32)
    w_cpu = torch.tensor(a["weights"])  # CPU tensor in the mapping path (photometric_factor.cpp:31-32)
    w_dev = w_cpu.to(dev)
    out = {}
    AtA, Atb, e = mod.photometric_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], jac0, R["code0"],
                                            R["mask1"], loc64, R["homo"], R["feat0"], R["feat1"], R["grad1"], lo, a["scale0"],
                                            cam, L, a["eps"], w_cpu)
    out.update(photo_AtA=AtA.cpu().numpy(), photo_Atb=Atb.cpu().numpy().reshape(-1), photo_err=e)
    out["photo_err_only"] = mod.photometric_error(R["R10"], R["t10"], R["bias0"], jac0, R["code0"], R["mask1"], loc64, R["homo"],
                                                  R["feat0"], R["feat1"], lo, a["scale0"], cam, L, a["eps"], w_cpu)
    AtA, Atb, e = mod.geometric_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], jac0, R["code0"],
                                          R["dpt1"], R["dgrad1"], R["basis1"], R["mask1"], loc32, R["homo"], a["scale0"],
                                          a["scale1"], cam, a["eps"], a["geo_loss"], a["geo_weight"])
    out.update(geo_AtA=AtA.cpu().numpy(), geo_Atb=Atb.cpu().numpy().reshape(-1), geo_err=e)
    out["geo_err_only"] = mod.geometric_error(R["R10"], R["t10"], R["bias0"], jac0, R["code0"], R["dpt1"], R["mask1"], loc32,
                                              R["homo"], a["scale0"], cam, a["eps"], a["geo_loss"], a["geo_weight"])
    mloc, mhomo, m2d, mdpts = T(ma["mloc"], dev, torch.int32), T(ma["mhomo"], dev), T(ma["m2d"], dev), T(ma["mdpts"], dev)
    AtA, Atb, e = mod.reprojection_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], jac0, R["code0"],
                                             mloc, mhomo, m2d, a["scale0"], cam, a["eps"], a["rep_loss"], a["rep_weight"])
    out.update(rep_AtA=AtA.cpu().numpy(), rep_Atb=Atb.cpu().numpy().reshape(-1), rep_err=e)
    out["rep_err_only"] = mod.reprojection_error(R["R10"], R["t10"], R["bias0"], jac0, R["code0"], mloc, mhomo, m2d, a["scale0"],
                                                 cam, a["eps"], a["rep_loss"], a["rep_weight"])
    sf, dp = T(ta["sfeat0"], dev), T(ta["dpts0"], dev)
    AtA, Atb, e = mod.tracker_photo_jac_error(R["R10"], R["t10"], R["mask1"], dp, R["homo"], sf, R["feat1"], R["grad1"], lo, cam, L,
                                              a["eps"], w_dev)
    out.update(trk_AtA=AtA.cpu().numpy(), trk_Atb=Atb.cpu().numpy().reshape(-1), trk_err=e)
    AtA, Atb, e = mod.tracker_photo_jac_error_with_scale(R["R10"], R["t10"], R["mask1"], dp, R["homo"], sf, R["feat1"], R["grad1"],
                                                         lo, cam, L, a["scale0"], a["eps"], w_dev)
    out.update(trks_AtA=AtA.cpu().numpy(), trks_Atb=Atb.cpu().numpy().reshape(-1), trks_err=e)
    out["trk_err_only"] = mod.tracker_photo_error(R["R10"], R["t10"], R["mask1"], dp, R["homo"], sf, R["feat1"], lo, cam, L,
                                                  a["eps"], w_dev)
    AtA, Atb, e = mod.tracker_reproj_jac_error(R["R10"], R["t10"], mdpts, mhomo, m2d, cam, a["eps"], a["rep_loss"], a["rep_weight"])
    out.update(trkrep_AtA=AtA.cpu().numpy(), trkrep_Atb=Atb.cpu().numpy().reshape(-1), trkrep_err=e)
    out["trkrep_err_only"] = mod.tracker_reproj_error(R["R10"], R["t10"], mdpts, mhomo, m2d, cam, a["eps"], a["rep_loss"],
                                                      a["rep_weight"])
    mh1, md1 = T(ma["mhomo1"], dev), T(ma["mdpts1"], dev)
    AtA, Atb, e = mod.tracker_match_geom_jac_error(R["R10"], R["t10"], mdpts, md1, mhomo, mh1, ma["mg_loss"], ma["mg_weight"])
    out.update(mg_AtA=AtA.cpu().numpy(), mg_Atb=Atb.cpu().numpy().reshape(-1), mg_err=e)
    AtA, Atb, e = mod.tracker_match_geom_jac_error_with_scale(R["R10"], R["t10"], mdpts, md1, mhomo, mh1, a["scale0"], ma["mg_loss"],
                                                              ma["mg_weight"])
    out.update(mgs_AtA=AtA.cpu().numpy(), mgs_Atb=Atb.cpu().numpy().reshape(-1), mgs_err=e)
    out["mg_err_only"] = mod.tracker_match_geom_error(R["R10"], R["t10"], mdpts, md1, mhomo, mh1, ma["mg_loss"], ma["mg_weight"])
    # mapping-side match geometry (all four robust losses) and the loop-closure form
    jac1 = T(np.ascontiguousarray(a["jac1"].T), dev).t()
    bias1, code1 = T(a["bias1"], dev), T(a["code1"], dev)
    mloc1 = T(ma["mloc1"], dev, torch.int32)
    for lt in helpers.MG_LOSSES:
        AtA, Atb, e = mod.match_geometry_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], R["bias0"], bias1, jac0, jac1,
                                                   R["code0"], code1, mhomo, mh1, mloc, mloc1, a["scale0"], a["scale1"], ma["mg_loss"],
                                                   ma["mg_weight"], lt)
        out.update({f"mmg_{lt}_AtA": AtA.cpu().numpy(), f"mmg_{lt}_Atb": Atb.cpu().numpy().reshape(-1), f"mmg_{lt}_err": e})
        out[f"mmg_{lt}_err_only"] = mod.match_geometry_error(R["R10"], R["t10"], R["bias0"], bias1, jac0, jac1, R["code0"], code1, mhomo,
                                                             mh1, mloc, mloc1, a["scale0"], a["scale1"], ma["mg_loss"], ma["mg_weight"],
                                                             lt)
    mu0, mu1 = T(ma["mud0"], dev), T(ma["mud1"], dev)
    AtA, Atb, e = mod.loop_mg_jac_error(R["R10"], R["t10"], R["R0"], R["t0"], R["R1"], R["t1"], mu0, mu1, mhomo, mh1, a["scale0"],
                                        a["scale1"], ma["mg_loss"], ma["mg_weight"])
    out.update(lmg_AtA=AtA.cpu().numpy(), lmg_Atb=Atb.cpu().numpy().reshape(-1), lmg_err=e)
    out["lmg_err_only"] = mod.loop_mg_error(R["R10"], R["t10"], mu0, mu1, mhomo, mh1, a["scale0"], a["scale1"], ma["mg_loss"],
                                            ma["mg_weight"])
    out["cam_pyramid"] = np.array(mod.camera_pyramid(cam, L), np.float32)
    out["sig"] = np.array([float(np.abs(a["feat0"]).sum()), float(np.abs(a["jac0"]).sum()), float(a["R10"].sum()),
                           float(ta["sfeat0"].sum()), float(ma["m2d"].sum())])
    return out


def main():
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])
    for name, far in helpers.GOLDEN_RUNS:
        if only and name not in only:
            continue
        if True:
            out = run_case(name, far)
            fn = os.path.join(outdir, f"{name}{'_far' if far else ''}.npz")
            np.savez_compressed(fn, **{k: np.asarray(v) for k, v in out.items()})
            print(fn, "photo_err", out["photo_err"], "geo_err", out["geo_err"], "rep_err", out["rep_err"], "trk_err", out["trk_err"])


if __name__ == "__main__":
    main()

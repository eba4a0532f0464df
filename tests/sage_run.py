"""Run the sm_100a kernels (through the C ABI) on a case: same keys as oracle_run.run_oracle."""
import numpy as np

import helpers
import sage_slam_b200 as sage
from sage_slam_b200 import ops


def run_sage(ctx, kfs):
    import torch

    a = helpers.case_args(kfs)
    ta = helpers.tracker_args(kfs, a)
    ma = helpers.match_args(kfs)
    d0, d1 = sage.DeviceKeyframe(ctx, kfs[0]), sage.DeviceKeyframe(ctx, kfs[1])
    out = {}
    A, b, e, n = ops.photometric_jac_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"],
                                                     a["scale0"], a["eps"], a["weights"])
    out.update(photo_AtA=A, photo_Atb=b, photo_err=e, photo_inl=n)
    out["photo_err_only"], _ = ops.photometric_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["code0"], a["scale0"], a["eps"],
                                                               a["weights"])
    A, b, e, n = ops.geometric_jac_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"],
                                                   a["code1"], a["scale0"], a["scale1"], a["eps"], a["geo_loss"], a["geo_weight"])
    out.update(geo_AtA=A, geo_Atb=b, geo_err=e, geo_inl=n)
    out["geo_err_only"], _ = ops.geometric_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["code0"], a["code1"], a["scale0"],
                                                           a["scale1"], a["eps"], a["geo_loss"], a["geo_weight"])
    A, b, e, n = ops.reprojection_jac_error_calculate(ctx, d0, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"],
                                                      a["scale0"], ma["mloc"], ma["mhomo"], ma["m2d"], a["eps"], a["rep_loss"],
                                                      a["rep_weight"])
    out.update(rep_AtA=A, rep_Atb=b, rep_err=e, rep_inl=n)
    out["rep_err_only"], _ = ops.reprojection_error_calculate(ctx, d0, a["R10"], a["t10"], a["code0"], a["scale0"], ma["mloc"],
                                                              ma["mhomo"], ma["m2d"], a["eps"], a["rep_loss"], a["rep_weight"])
    # tracker: feed the SAME pre-sampled tensors the reference/oracle got, plus check our own presampler
    smp = ops.TrackerSamples(ctx, d0, a["code0"], a["scale0"])
    out["presample_feats"] = smp.feats.cpu().numpy()
    out["presample_dpts"] = smp.dpts.cpu().numpy()
    dev = smp.feats.device
    smp.feats = torch.from_numpy(ta["sfeat0"]).to(dev)
    smp.dpts = torch.from_numpy(ta["dpts0"]).to(dev)
    torch.cuda.synchronize()
    A, b, e, n = ops.tracker_photo_jac_error_calculate(ctx, d1, a["R10"], a["t10"], smp, a["eps"], a["weights"])
    out.update(trk_AtA=A, trk_Atb=b, trk_err=e)
    A, b, e, n = ops.tracker_photo_jac_error_calculate(ctx, d1, a["R10"], a["t10"], smp, a["eps"], a["weights"], scale_0=a["scale0"])
    out.update(trks_AtA=A, trks_Atb=b, trks_err=e)
    out["trk_err_only"], _ = ops.tracker_photo_error_calculate(ctx, d1, a["R10"], a["t10"], smp, a["eps"], a["weights"])
    A, b, e, n = ops.tracker_reproj_jac_error_calculate(ctx, a["cam"], a["R10"], a["t10"], ma["mdpts"], ma["mhomo"], ma["m2d"],
                                                        a["eps"], a["rep_loss"], a["rep_weight"])
    out.update(trkrep_AtA=A, trkrep_Atb=b, trkrep_err=e)
    out["trkrep_err_only"], _ = ops.tracker_reproj_error_calculate(ctx, a["cam"], a["R10"], a["t10"], ma["mdpts"], ma["mhomo"],
                                                                   ma["m2d"], a["eps"], a["rep_loss"], a["rep_weight"])
    A, b, e = ops.tracker_match_geom_jac_error_calculate(ctx, a["R10"], a["t10"], ma["mdpts"], ma["mdpts1"], ma["mhomo"], ma["mhomo1"],
                                                         ma["mg_loss"], ma["mg_weight"])
    out.update(mg_AtA=A, mg_Atb=b, mg_err=e)
    A, b, e = ops.tracker_match_geom_jac_error_calculate(ctx, a["R10"], a["t10"], ma["mdpts"], ma["mdpts1"], ma["mhomo"], ma["mhomo1"],
                                                         ma["mg_loss"], ma["mg_weight"], scale_0=a["scale0"])
    out.update(mgs_AtA=A, mgs_Atb=b, mgs_err=e)
    out["mg_err_only"] = ops.tracker_match_geom_error_calculate(ctx, a["R10"], a["t10"], ma["mdpts"], ma["mdpts1"], ma["mhomo"],
                                                                ma["mhomo1"], ma["mg_loss"], ma["mg_weight"])
    for lt in helpers.MG_LOSSES:
        A, b, e = ops.match_geometry_jac_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["code0"],
                                                         a["code1"], ma["mhomo"], ma["mhomo1"], ma["mloc"], ma["mloc1"], a["scale0"],
                                                         a["scale1"], ma["mg_loss"], ma["mg_weight"], lt)
        out.update({f"mmg_{lt}_AtA": A, f"mmg_{lt}_Atb": b, f"mmg_{lt}_err": e})
        out[f"mmg_{lt}_err_only"] = ops.match_geometry_error_calculate(ctx, d0, d1, a["R10"], a["t10"], a["code0"], a["code1"], ma["mhomo"],
                                                                       ma["mhomo1"], ma["mloc"], ma["mloc1"], a["scale0"], a["scale1"],
                                                                       ma["mg_loss"], ma["mg_weight"], lt)
    A, b, e = ops.loop_mg_jac_error_calculate(ctx, a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], ma["mud0"], ma["mud1"],
                                              ma["mhomo"], ma["mhomo1"], a["scale0"], a["scale1"], ma["mg_loss"], ma["mg_weight"])
    out.update(lmg_AtA=A, lmg_Atb=b, lmg_err=e)
    out["lmg_err_only"] = ops.loop_mg_error_calculate(ctx, a["R10"], a["t10"], ma["mud0"], ma["mud1"], ma["mhomo"], ma["mhomo1"],
                                                      a["scale0"], a["scale1"], ma["mg_loss"], ma["mg_weight"])
    out["cam_pyramid"] = d0.cameras()[0]
    out["ref_sfeat0"] = ta["sfeat0"]
    out["ref_dpts0"] = ta["dpts0"]
    return out

"""Run the CPU oracle on a case: one dict with the same keys oracle/make_golden.py stores."""
import numpy as np

import helpers
import oracle as O


def run_oracle(kfs, dtype=np.float32):
    a = helpers.case_args(kfs)
    ta = helpers.tracker_args(kfs, a)
    ma = helpers.match_args(kfs)
    out = {}
    A, b, e, n = O.photometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                         a["mask1"], a["loc1d"], a["homo"], a["feat0"], a["feat1"], a["grad1"], a["level_offsets"],
                                         a["scale0"], a["cams"], a["eps"], a["weights"], dtype=dtype)
    out.update(photo_AtA=A, photo_Atb=b, photo_err=e, photo_inl=n)
    out["photo_err_only"], _ = O.photometric_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], a["mask1"], a["loc1d"],
                                                   a["homo"], a["feat0"], a["feat1"], a["level_offsets"], a["scale0"], a["cams"],
                                                   a["eps"], a["weights"], dtype=dtype)
    A, b, e, n = O.geometric_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                       a["dpt1"], a["dgrad1"], a["basis1"], a["mask1"], a["loc1d"], a["homo"], a["scale0"],
                                       a["scale1"], a["cam"], a["eps"], a["geo_loss"], a["geo_weight"], dtype=dtype)
    out.update(geo_AtA=A, geo_Atb=b, geo_err=e, geo_inl=n)
    out["geo_err_only"], _ = O.geometric_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], a["dpt1"], a["mask1"],
                                               a["loc1d"], a["homo"], a["scale0"], a["cam"], a["eps"], a["geo_loss"],
                                               a["geo_weight"], dtype=dtype)
    A, b, e, n = O.reprojection_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["jac0"], a["code0"],
                                          ma["mloc"], ma["mhomo"], ma["m2d"], a["scale0"], a["cam"], a["eps"], a["rep_loss"],
                                          a["rep_weight"], dtype=dtype)
    out.update(rep_AtA=A, rep_Atb=b, rep_err=e, rep_inl=n)
    out["rep_err_only"], _ = O.reprojection_error(a["R10"], a["t10"], a["bias0"], a["jac0"], a["code0"], ma["mloc"], ma["mhomo"],
                                                  ma["m2d"], a["scale0"], a["cam"], a["eps"], a["rep_loss"], a["rep_weight"],
                                                  dtype=dtype)
    A, b, e, n = O.tracker_photo_jac_error(a["R10"], a["t10"], a["mask1"], ta["dpts0"], a["homo"], ta["sfeat0"], a["feat1"],
                                           a["grad1"], a["level_offsets"], a["cams"], a["eps"], a["weights"], dtype=dtype)
    out.update(trk_AtA=A, trk_Atb=b, trk_err=e)
    A, b, e, n = O.tracker_photo_jac_error(a["R10"], a["t10"], a["mask1"], ta["dpts0"], a["homo"], ta["sfeat0"], a["feat1"],
                                           a["grad1"], a["level_offsets"], a["cams"], a["eps"], a["weights"], scale0=a["scale0"],
                                           dtype=dtype)
    out.update(trks_AtA=A, trks_Atb=b, trks_err=e)
    out["trk_err_only"], _ = O.tracker_photo_error(a["R10"], a["t10"], a["mask1"], ta["dpts0"], a["homo"], ta["sfeat0"], a["feat1"],
                                                   a["level_offsets"], a["cams"], a["eps"], a["weights"], dtype=dtype)
    A, b, e, n = O.tracker_reproj_jac_error(a["R10"], a["t10"], ma["mdpts"], ma["mhomo"], ma["m2d"], a["cam"], a["eps"],
                                            a["rep_loss"], a["rep_weight"], dtype=dtype)
    out.update(trkrep_AtA=A, trkrep_Atb=b, trkrep_err=e)
    out["trkrep_err_only"], _ = O.tracker_reproj_error(a["R10"], a["t10"], ma["mdpts"], ma["mhomo"], ma["m2d"], a["cam"], a["eps"],
                                                       a["rep_loss"], a["rep_weight"], dtype=dtype)
    A, b, e = O.tracker_match_geom_jac_error(a["R10"], a["t10"], ma["mdpts"], ma["mdpts1"], ma["mhomo"], ma["mhomo1"], ma["mg_loss"],
                                             ma["mg_weight"], dtype=dtype)
    out.update(mg_AtA=A, mg_Atb=b, mg_err=e)
    A, b, e = O.tracker_match_geom_jac_error(a["R10"], a["t10"], ma["mdpts"], ma["mdpts1"], ma["mhomo"], ma["mhomo1"], ma["mg_loss"],
                                             ma["mg_weight"], scale0=a["scale0"], dtype=dtype)
    out.update(mgs_AtA=A, mgs_Atb=b, mgs_err=e)
    out["mg_err_only"] = O.tracker_match_geom_error(a["R10"], a["t10"], ma["mdpts"], ma["mdpts1"], ma["mhomo"], ma["mhomo1"],
                                                    ma["mg_loss"], ma["mg_weight"], dtype=dtype)
    for lt in helpers.MG_LOSSES:
        A, b, e = O.match_geometry_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], a["bias0"], a["bias1"], a["jac0"],
                                             a["jac1"], a["code0"], a["code1"], ma["mhomo"], ma["mhomo1"], ma["mloc"], ma["mloc1"],
                                             a["scale0"], a["scale1"], ma["mg_loss"], ma["mg_weight"], lt, dtype=dtype)
        out.update({f"mmg_{lt}_AtA": A, f"mmg_{lt}_Atb": b, f"mmg_{lt}_err": e})
        out[f"mmg_{lt}_err_only"] = O.match_geometry_error(a["R10"], a["t10"], a["bias0"], a["bias1"], a["jac0"], a["jac1"], a["code0"],
                                                           a["code1"], ma["mhomo"], ma["mhomo1"], ma["mloc"], ma["mloc1"], a["scale0"],
                                                           a["scale1"], ma["mg_loss"], ma["mg_weight"], lt, dtype=dtype)
    A, b, e = O.loop_mg_jac_error(a["R10"], a["t10"], a["R0"], a["t0"], a["R1"], a["t1"], ma["mud0"], ma["mud1"], ma["mhomo"],
                                  ma["mhomo1"], a["scale0"], a["scale1"], ma["mg_loss"], ma["mg_weight"], dtype=dtype)
    out.update(lmg_AtA=A, lmg_Atb=b, lmg_err=e)
    out["lmg_err_only"] = O.loop_mg_error(a["R10"], a["t10"], ma["mud0"], ma["mud1"], ma["mhomo"], ma["mhomo1"], a["scale0"],
                                          a["scale1"], ma["mg_loss"], ma["mg_weight"], dtype=dtype)
    out["cam_pyramid"] = O.camera_pyramid(a["cam"], a["L"])
    out["sig"] = np.array([float(np.abs(a["feat0"]).sum()), float(np.abs(a["jac0"]).sum()), float(a["R10"].sum()),
                           float(ta["sfeat0"].sum()), float(ma["m2d"].sum())])
    return out

"""ctypes binding of include/sage_ba.h (libsage_ba.so).  No fallbacks: if the CUDA library is missing
or no GPU is present the product path raises -- nothing here ever routes through oracle/."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAGE_BA_LIB", os.path.join(HERE, "lib", "libsage_ba.so"))
MAX_LEVELS = 8
PROF_KINDS = ["photo_jac", "geo_jac", "reproj_jac", "photo_err", "geo_err", "reproj_err", "depth_prep", "assemble", "solve", "comm"]
HOST, DEVICE = 0, 1

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
vp = C.c_void_p


class Camera(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("fx", "fy", "u0", "v0", "width", "height")]


class KeyframeDesc(C.Structure):
    _fields_ = [("memory", C.c_int), ("height", C.c_int), ("width", C.c_int), ("levels", C.c_int),
                ("feat_channels", C.c_int), ("code_size", C.c_int), ("camera", Camera),
                ("feat_map", vp), ("feat_map_pyramid", vp), ("feat_map_grad_pyramid", vp), ("dpt_map_bias", vp), ("dpt_jac_code", vp),
                ("jac_stride_row", C.c_long), ("jac_stride_col", C.c_long), ("video_mask", vp),
                ("sampled_locations_1d", vp), ("sampled_locations_homo", vp), ("num_samples", C.c_int), ("borrow_depth", C.c_int)]


class TrackerConfig(C.Structure):
    _fields_ = [("max_num_iters", C.c_int), ("init_damp", C.c_float), ("min_damp", C.c_float), ("max_damp", C.c_float),
                ("damp_dec_factor", C.c_float), ("damp_inc_factor", C.c_float),
                ("jac_update_err_inc_threshold", C.c_float), ("min_grad_thresh", C.c_float),
                ("min_param_inc_thresh", C.c_float), ("dpt_eps", C.c_float), ("photo_weights", C.c_float * MAX_LEVELS),
                ("use_photo", C.c_int), ("use_reproj", C.c_int), ("reproj_loss_param", C.c_float),
                ("reproj_weight", C.c_float), ("use_match_geom", C.c_int), ("match_geom_loss_param", C.c_float),
                ("match_geom_weight", C.c_float)]


class TrackerReport(C.Structure):
    _fields_ = [("iterations", C.c_int), ("jacobian_evals", C.c_int), ("error_evals", C.c_int),
                ("final_error", C.c_float), ("final_damp", C.c_float)]


class LMOptions(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("init_damp", C.c_double), ("min_damp", C.c_double), ("max_damp", C.c_double),
                ("damp_dec_factor", C.c_double), ("damp_inc_factor", C.c_double), ("min_rel_decrease", C.c_double),
                ("max_trials", C.c_int)]


class LMReport(C.Structure):
    _fields_ = [("iterations", C.c_int), ("linearizations", C.c_int), ("evaluations", C.c_int), ("accepted", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_damp", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, vp, C.c_size_t, vp)
LM_JAC_FN = C.CFUNCTYPE(None, vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                        C.POINTER(C.c_float))
LM_ERR_FN = C.CFUNCTYPE(C.c_float, vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float)

# name -> (restype, argtypes); every symbol include/sage_ba.h declares
F = C.c_float
SIGNATURES = {
    "sage_ba_create": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
    "sage_ba_destroy": (None, [vp]),
    "sage_ba_last_error": (C.c_char_p, [vp]),
    "sage_ba_version": (C.c_char_p, []),
    "sage_ba_launch_count": (C.c_long, [vp]),
    "sage_ba_synchronize": (C.c_int, [vp]),
    "sage_ba_stream": (vp, [vp]),
    "sage_ba_set_geometric_tcgen05": (C.c_int, [C.c_int]),
    "sage_ba_keyframe_create": (C.c_int, [vp, C.POINTER(KeyframeDesc), C.POINTER(vp)]),
    "sage_ba_keyframe_destroy": (None, [vp, vp]),
    "sage_ba_keyframe_set_bias": (C.c_int, [vp, vp, vp, C.c_int]),
    "sage_ba_keyframe_cameras": (C.c_int, [vp, C.POINTER(Camera), c_int_p]),
    "sage_ba_photometric_jac_error": (C.c_int, [vp, vp, vp] + [vp] * 7 + [F, F, vp, vp, vp, vp, vp]),
    "sage_ba_photometric_error": (C.c_int, [vp, vp, vp, vp, vp, vp, F, F, vp, vp, vp]),
    "sage_ba_tracker_photo_jac_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, F, F, vp, vp, vp, vp, vp]),
    "sage_ba_tracker_photo_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, F, vp, vp, vp]),
    "sage_ba_tracker_presample": (C.c_int, [vp, vp, vp, F, vp, vp, vp]),
    "sage_ba_geometric_jac_error": (C.c_int, [vp, vp, vp] + [vp] * 8 + [F, F, F, F, F, vp, vp, vp, vp]),
    "sage_ba_geometric_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, F, F, F, F, F, vp, vp]),
    "sage_ba_reprojection_jac_error": (C.c_int, [vp, vp] + [vp] * 7 + [F, vp, vp, vp, C.c_int, F, F, F, vp, vp, vp, vp]),
    "sage_ba_reprojection_error": (C.c_int, [vp, vp, vp, vp, vp, F, vp, vp, vp, C.c_int, F, F, F, vp, vp]),
    "sage_ba_tracker_reproj_jac_error": (C.c_int, [vp, C.POINTER(Camera), vp, vp, vp, vp, vp, C.c_int, F, F, F, vp, vp, vp, vp]),
    "sage_ba_tracker_reproj_error": (C.c_int, [vp, C.POINTER(Camera), vp, vp, vp, vp, vp, C.c_int, F, F, F, vp, vp]),
    "sage_ba_tracker_match_geom_jac_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, F, F, F, vp, vp, vp]),
    "sage_ba_tracker_match_geom_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, F, F, vp]),
    "sage_ba_match_geometry_jac_error": (C.c_int, [vp, vp, vp] + [vp] * 8 + [F, F, vp, vp, vp, vp, C.c_int, F, F, C.c_int, vp, vp, vp]),
    "sage_ba_match_geometry_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, F, F, vp, vp, vp, vp, C.c_int, F, F, C.c_int, vp]),
    "sage_ba_loop_mg_jac_error": (C.c_int, [vp] + [vp] * 10 + [C.c_int, F, F, F, F, vp, vp, vp]),
    "sage_ba_loop_mg_error": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, F, F, F, F, vp]),
    "sage_ba_cycle_match": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, F, vp, vp, vp, c_int_p, c_float_p]),
    "sage_ba_track_frame": (C.c_int, [vp, vp, vp, vp, C.POINTER(TrackerConfig), vp, vp, vp, vp, vp, vp, vp, C.c_int,
                                      C.POINTER(TrackerReport)]),
    "sage_ba_track_new_frame": (C.c_int, [vp, vp, vp, vp, F, C.POINTER(TrackerConfig), vp, vp, vp, vp, vp, C.c_int,
                                          C.POINTER(TrackerReport)]),
    "sage_ba_tracker_solve": (C.c_int, [vp, vp, C.c_int, F, vp]),
    "sage_ba_se3_exp": (C.c_int, [vp, vp, vp, vp]),
    "sage_ba_tracker_lm_callbacks": (C.c_int, [C.c_int, C.POINTER(TrackerConfig), vp, vp, vp, LM_JAC_FN, LM_ERR_FN, vp,
                                               C.POINTER(TrackerReport)]),
    "sage_ba_problem_create": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "sage_ba_problem_destroy": (None, [vp]),
    "sage_ba_problem_add_photometric": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "sage_ba_problem_add_geometric": (C.c_int, [vp, C.c_int, C.c_int, F, F]),
    "sage_ba_problem_add_reprojection": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, F, F]),
    "sage_ba_problem_add_code_prior": (C.c_int, [vp, C.c_int, vp, F]),
    "sage_ba_problem_add_scale_prior": (C.c_int, [vp, C.c_int, F, F]),
    "sage_ba_problem_fix": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
    "sage_ba_problem_set_solver": (C.c_int, [vp, C.c_int]),
    "sage_ba_problem_set_shard": (C.c_int, [vp, C.c_int, C.c_int]),
    "sage_ba_shard_plan": (C.c_int, [C.c_int, c_int_p, c_int_p, C.c_int, c_int_p]),
    "sage_ba_problem_set_relinearize_always": (C.c_int, [vp, C.c_int]),
    "sage_ba_problem_set_deterministic": (C.c_int, [vp, C.c_int]),
    "sage_ba_problem_factor_offsets": (C.c_int, [vp, c_int_p, c_int_p, c_int_p]),
    "sage_ba_problem_solver_info": (C.c_int, [vp, c_int_p, c_int_p, c_int_p]),
    "sage_ba_problem_solver_trace": (C.c_int, [vp, vp, c_int_p]),
    "sage_ba_nccl_unique_id": (C.c_int, [vp]),
    "sage_ba_comm_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "sage_ba_comm_wrap": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "sage_ba_comm_destroy": (None, [vp]),
    "sage_ba_problem_set_comm": (C.c_int, [vp, vp]),
    "sage_ba_problem_exchange": (C.c_int, [vp, C.c_int]),
    "sage_ba_problem_set_state": (C.c_int, [vp, vp, vp, vp, F]),
    "sage_ba_problem_get_state": (C.c_int, [vp, vp, vp, vp]),
    "sage_ba_problem_update_map": (C.c_int, [vp, vp, vp, vp, vp, C.c_int]),
    "sage_ba_problem_dim": (C.c_int, [vp]),
    "sage_ba_problem_num_factors": (C.c_int, [vp]),
    "sage_ba_problem_num_residuals": (C.c_long, [vp]),
    "sage_ba_problem_factor_buffer": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "sage_ba_problem_cost_buffer": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "sage_ba_problem_linearize": (C.c_int, [vp]),
    "sage_ba_problem_assemble": (C.c_int, [vp, vp, vp, c_double_p]),
    "sage_ba_problem_solve": (C.c_int, [vp, C.c_double, vp]),
    "sage_ba_problem_evaluate": (C.c_int, [vp, C.c_int]),
    "sage_ba_problem_cost": (C.c_int, [vp, C.c_int, c_double_p]),
    "sage_ba_problem_accept": (C.c_int, [vp]),
    "sage_ba_problem_set_allreduce": (C.c_int, [vp, ALLREDUCE_FN, vp]),
    "sage_ba_problem_profile": (C.c_int, [vp, C.c_int]),
    "sage_ba_problem_profile_read": (C.c_int, [vp, c_double_p, C.POINTER(C.c_long), C.c_int]),
    "sage_ba_problem_shard_counts": (C.c_int, [vp, c_int_p, c_int_p, c_int_p]),
    "sage_ba_problem_lm_step": (C.c_int, [vp, c_double_p, C.c_double, C.c_double, C.c_double, C.c_double, c_double_p, c_double_p, c_int_p]),
    "sage_ba_problem_lm": (C.c_int, [vp, C.POINTER(LMOptions), C.POINTER(LMReport)]),
}

_lib = None


def load():
    """dlopen libsage_ba.so and attach signatures; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python sage-slam_b200/build.py` (nvcc, sm_100a). "
                               "There is no CPU fallback for the factor kernels.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

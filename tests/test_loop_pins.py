"""CPU: the tracker LM loop bookkeeping pinned by the reference's OWN loops.

tests/golden/loop_pins.npz holds, for 12 synthetic problems, every cost evaluation CameraTracker::TrackNewFrame (6-DoF,
core/system/camera_tracker.cpp:1156-1279) and CameraTracker::TrackFrame (7-DoF, :1479-1630) make -- the loops, UpdateVariables and
LMConvergence extracted verbatim from the reference at build time and run over a reprojection cost (oracle/build_loop_ref.py,
oracle/loop_pins.cpp, oracle/make_golden_loop.py).  oracle.tracker_lm / tracker_lm7 -- the restatements the GPU tracker
(sage_ba_track_new_frame / sage_ba_track_frame) is tested against in tests/test_gpu_parity.py -- must make the same evaluations
in the same order at the same states: which steps are tried, rejected and retried with more damping, when the Jacobian is
skipped, and where the loop stops (convergence test, maximum damping, iteration limit)."""
import os

import numpy as np
import pytest

import helpers
import make_golden_loop as G
import oracle as O

PINS = os.path.join(helpers.ROOT, "tests", "golden", "loop_pins.npz")
CASES = [c["name"] for c in G.make_cases()]
# the reference solves the damped 6x6 / 7x7 system with Eigen's float colPivHouseholderQr, the restatement in fp64: steps agree to
# ~1e-5 of their length (condition ~1e6 at damping 1e-4 x fp32), the evaluation ORDER must agree exactly
TOL_STATE, TOL_ERR = 5e-5, 2e-4


def _replay(case, pins):
    n = case["name"]
    keys = [str(k) for k in pins["opt_keys"]]
    opt = dict(zip(keys, pins[n + "/opt"]))
    jac_fn, err_fn = G.reprojection_cost(pins[n + "/pts"], tuple(pins["cam"]), case["dof"])
    log = []

    def jac(R, t, s=1.0):
        A, b, e = jac_fn(R, t, s) if case["dof"] == 7 else jac_fn(R, t)
        log.append(("J", np.asarray(R, np.float64).reshape(-1), np.asarray(t, np.float64).reshape(-1), float(s), float(e)))
        return A, b, e

    def err(R, t, s=1.0):
        e = err_fn(R, t, s) if case["dof"] == 7 else err_fn(R, t)
        log.append(("E", np.asarray(R, np.float64).reshape(-1), np.asarray(t, np.float64).reshape(-1), float(s), float(e)))
        return e

    kw = dict(init_damp=opt["init_damp"], min_damp=opt["min_damp"], max_damp=opt["max_damp"], damp_dec=opt["damp_dec"], damp_inc=opt["damp_inc"],
              max_iters=int(opt["max_iters"]), jac_thresh=opt["jac_thresh"], min_grad=opt["min_grad"], min_param_inc=opt["min_param_inc"])
    if case["dof"] == 6:
        R, t, e, _ = O.tracker_lm(jac, err, pins[n + "/R0"], pins[n + "/t0"], **kw)
        s = 1.0
    else:
        R, t, s, e, _ = O.tracker_lm7(jac, err, pins[n + "/R0"], pins[n + "/t0"], float(pins[n + "/s0"]), **kw)
    return log, (np.asarray(R, np.float64).reshape(-1), np.asarray(t, np.float64).reshape(-1), float(s), float(e))


def test_goldens_cover_every_branch_of_the_loops():
    pins = np.load(PINS)
    kinds = {n: str(pins[n + "/kinds"]) for n in CASES}
    assert len(kinds) == 12 and all(k.startswith("JE") for k in kinds.values())
    assert any("EE" in k for k in kinds.values())                      # a rejected step retried with more damping
    assert any(k.endswith("EEEE") for k in kinds.values())             # damping raised until the maximum: termination there
    assert any("JEEJ" in k or "EJEE" in k for k in kinds.values())     # rejection in the middle of a run
    it = {n: int(pins[n + "/final"][-1]) for n in CASES}
    assert it["dof6_few_iters"] < 3 or it["dof7_few_iters"] == 3       # the iteration limit
    # Jacobian skipped: more iterations than Jacobian evaluations
    assert any(it[n] > kinds[n].count("J") for n in CASES)


@pytest.mark.parametrize("name", CASES)
def test_restated_loops_make_the_reference_loops_evaluations(name):
    pins = np.load(PINS)
    case = next(c for c in G.make_cases() if c["name"] == name)
    np.testing.assert_array_equal(case["pts"], pins[name + "/pts"])  # the generator is deterministic: the fixture is reproducible
    log, final = _replay(case, pins)
    want_kinds, want = str(pins[name + "/kinds"]), pins[name + "/log"]
    got_kinds = "".join(k for k, *_ in log)
    assert got_kinds == want_kinds, f"{name}: evaluation order {got_kinds} != reference {want_kinds}"
    for i, (k, R, t, s, e) in enumerate(log):
        np.testing.assert_allclose(R, want[i, :9], atol=TOL_STATE, err_msg=f"{name} evaluation {i} ({k}): rotation")
        np.testing.assert_allclose(t, want[i, 9:12], atol=TOL_STATE, err_msg=f"{name} evaluation {i} ({k}): translation")
        assert abs(s - want[i, 12]) <= TOL_STATE, (name, i, s, want[i, 12])
        assert abs(e - want[i, 13]) <= TOL_ERR * abs(want[i, 13]), (name, i, e, want[i, 13])
    R, t, s, e = final
    fin = pins[name + "/final"]
    np.testing.assert_allclose(R, fin[:9], atol=TOL_STATE)
    np.testing.assert_allclose(t, fin[9:12], atol=TOL_STATE)
    assert abs(s - fin[12]) <= TOL_STATE and abs(e - fin[13]) <= TOL_ERR * abs(fin[13])


def _replay_product(case, pins):
    """The library's own loop (csrc/tracker.cu lm_loop, the code behind sage_ba_track_new_frame / sage_ba_track_frame) over the same
    cost through sage_ba_tracker_lm_callbacks: host code, no GPU."""
    import ctypes as C

    import sage_slam_b200 as sage
    from sage_slam_b200 import capi

    lib = sage.capi.load()
    n, dof = case["name"], case["dof"]
    opt = dict(zip([str(k) for k in pins["opt_keys"]], pins[n + "/opt"]))
    jac_fn, err_fn = G.reprojection_cost(pins[n + "/pts"], tuple(pins["cam"]), dof)
    log = []

    def state(Rp, tp, s):
        R = np.array([Rp[i] for i in range(9)], np.float32).reshape(3, 3)
        t = np.array([tp[i] for i in range(3)], np.float32)
        return R, t, np.float32(s)

    def jac(_user, Rp, tp, s, Ap, bp, ep):
        R, t, s = state(Rp, tp, s)
        A, b, e = jac_fn(R, t, s) if dof == 7 else jac_fn(R, t)
        for i, v in enumerate(np.asarray(A, np.float32).reshape(-1)):
            Ap[i] = v
        for i, v in enumerate(np.asarray(b, np.float32).reshape(-1)):
            bp[i] = v
        ep[0] = e
        log.append(("J", R.astype(np.float64).reshape(-1), t.astype(np.float64), float(s), float(e)))

    def err(_user, Rp, tp, s):
        R, t, s = state(Rp, tp, s)
        e = float(err_fn(R, t, s) if dof == 7 else err_fn(R, t))
        log.append(("E", R.astype(np.float64).reshape(-1), t.astype(np.float64), float(s), e))
        return e

    cfg = capi.TrackerConfig()
    cfg.max_num_iters = int(opt["max_iters"])
    cfg.init_damp, cfg.min_damp, cfg.max_damp = opt["init_damp"], opt["min_damp"], opt["max_damp"]
    cfg.damp_dec_factor, cfg.damp_inc_factor = opt["damp_dec"], opt["damp_inc"]
    cfg.jac_update_err_inc_threshold, cfg.min_grad_thresh, cfg.min_param_inc_thresh = opt["jac_thresh"], opt["min_grad"], opt["min_param_inc"]
    R = np.ascontiguousarray(pins[n + "/R0"], np.float32).copy()
    t = np.ascontiguousarray(pins[n + "/t0"], np.float32).reshape(-1).copy()
    s = C.c_float(float(pins[n + "/s0"]))
    rep = capi.TrackerReport()
    cb_j, cb_e = capi.LM_JAC_FN(jac), capi.LM_ERR_FN(err)
    rc = lib.sage_ba_tracker_lm_callbacks(dof, C.byref(cfg), R.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                          C.cast(C.byref(s), C.c_void_p), cb_j, cb_e, None, C.byref(rep))
    assert rc == 0
    return log, (R.astype(np.float64).reshape(-1), t.astype(np.float64), float(s.value) if dof == 7 else 1.0, float(rep.final_error)), rep


@pytest.mark.parametrize("name", CASES)
def test_library_loop_makes_the_reference_loops_evaluations(name):
    """Row a7 for the PRODUCT: the C++ loop of the library, driven with the synthetic cost, against the reference's own loop."""
    pins = np.load(PINS)
    case = next(c for c in G.make_cases() if c["name"] == name)
    log, final, rep = _replay_product(case, pins)
    want_kinds, want, fin = str(pins[name + "/kinds"]), pins[name + "/log"], pins[name + "/final"]
    got_kinds = "".join(k for k, *_ in log)
    assert got_kinds == want_kinds, f"{name}: evaluation order {got_kinds} != reference {want_kinds}"
    assert rep.jacobian_evals == want_kinds.count("J") and rep.error_evals == want_kinds.count("E")
    assert rep.iterations == int(fin[-1]), (rep.iterations, fin[-1])
    for i, (k, R, t, s, e) in enumerate(log):
        np.testing.assert_allclose(R, want[i, :9], atol=TOL_STATE, err_msg=f"{name} evaluation {i} ({k}): rotation")
        np.testing.assert_allclose(t, want[i, 9:12], atol=TOL_STATE, err_msg=f"{name} evaluation {i} ({k}): translation")
        assert abs(s - want[i, 12]) <= TOL_STATE, (name, i, s, want[i, 12])
        assert abs(e - want[i, 13]) <= TOL_ERR * abs(want[i, 13]), (name, i, e, want[i, 13])
    R, t, s, e = final
    np.testing.assert_allclose(R, fin[:9], atol=TOL_STATE)
    np.testing.assert_allclose(t, fin[9:12], atol=TOL_STATE)
    assert abs(s - fin[12]) <= TOL_STATE and abs(e - fin[13]) <= TOL_ERR * abs(fin[13])


@pytest.mark.parametrize("case", G.update_depth_cases(), ids=lambda c: c["name"])
def test_update_depth_matches_the_references_own_function(case):
    """Row f4 (Mapper::UpdateMap's depth write-back): oracle.update_depth -- the checker of sage_ba_problem_update_map and of the
    mapper adapters (tests/test_mapper.py, tests/test_gpu_solver.py) -- against UpdateDepth<float> itself
    (core/mapping/mapping_utils.h:216-222, extracted at build time, run with libtorch on the CPU)."""
    pins = np.load(PINS)
    got = O.update_depth(case["bias"], case["jac"], case["code"], case["scale"])
    want = pins[case["name"] + "/dpt_map"]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=5e-7)  # one fp32 GEMV: summation order is the only freedom (1 ulp where bias + jac.code cancels)

"""Generate tests/golden/loop_pins.npz: every cost evaluation the reference's OWN tracker LM loops make on synthetic problems.

    python oracle/make_golden_loop.py          (only where /root/reference exists; see oracle/build_loop_ref.py)

Each case is a point cloud seen by a pinhole camera, a perturbed start pose (and depth scale for the 7-DoF loop) and one set of
LM options (the reference's run-time flags, system/configs/slam_run.flags:17-23, and variations that reach the other branches:
rejected steps with growing damping, the Jacobian-skip branch, termination at max damping and at the iteration limit).  The log
is the sequence of (kind J|E, pose, scale, error) in call order plus the final state; tests/test_loop_pins.py replays the same
cases through oracle.tracker_lm / tracker_lm7 with `reprojection_cost` below and compares evaluation by evaluation.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "..", "tests", "golden", "loop_pins.npz")
CAM = (160.0, 150.0, 80.0, 64.0)  # fx fy cx cy
FLAGS = dict(init_damp=1e-4, min_damp=1e-6, max_damp=1e-2, damp_inc=100.0, damp_dec=10.0, jac_thresh=1e-2, min_grad=1e-4, min_param_inc=1e-2,
             max_iters=40)  # slam_run.flags:17-23


def rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


WZ = {6: 0.0, 7: 40.0}  # weight of the depth residual: the 7-DoF cases need it to fix the scale gauge (pixels per unit of depth)


def reprojection_cost(pts, cam, dof):
    """The cost of oracle/loop_pins.cpp restated: x = R (s p) + t, r = [pi(x) - uv ; wz (x.z - d)], float64 arithmetic rounded to
    float32.  Returns (jac_fn, err_fn) in the signatures oracle.tracker_lm (dof 6) / tracker_lm7 (dof 7) expect."""
    fx, fy, cx, cy = cam
    wz = WZ[dof]
    P, UV, Dz = pts[:, :3].astype(np.float64), pts[:, 3:5].astype(np.float64), pts[:, 5].astype(np.float64)

    def terms(R, t, s):
        R, t = np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3)
        Rp = P @ R.T
        x = s * Rp + t
        iz = 1.0 / x[:, 2]
        r = np.stack([fx * x[:, 0] * iz + cx - UV[:, 0], fy * x[:, 1] * iz + cy - UV[:, 1], wz * (x[:, 2] - Dz)], 1)
        return Rp, x, iz, r

    def err(R, t, s=1.0):
        r = terms(R, t, float(s))[3]
        return np.float32(np.sum(r * r) / len(P))

    def jac(R, t, s=1.0):
        Rp, x, iz, r = terms(R, t, float(s))
        n = len(P)
        Pj = np.zeros((n, 3, 3))
        Pj[:, 2, 2] = wz
        Pj[:, 0, 0], Pj[:, 0, 2] = fx * iz, -fx * x[:, 0] * iz * iz
        Pj[:, 1, 1], Pj[:, 1, 2] = fy * iz, -fy * x[:, 1] * iz * iz
        D = np.zeros((n, 3, 7))
        D[:, 0, 0] = D[:, 1, 1] = D[:, 2, 2] = 1.0
        D[:, 0, 4], D[:, 0, 5] = x[:, 2], -x[:, 1]
        D[:, 1, 3], D[:, 1, 5] = -x[:, 2], x[:, 0]
        D[:, 2, 3], D[:, 2, 4] = x[:, 1], -x[:, 0]
        D[:, :, 6] = Rp
        J = (Pj @ D)[:, :, :dof]
        A = np.einsum("nai,naj->ij", J, J) / n
        b = -np.einsum("nai,na->i", J, r) / n
        return A.astype(np.float32), b.astype(np.float32), np.float32(np.sum(r * r) / n)

    if dof == 6:
        return (lambda R, t: jac(R, t)), (lambda R, t: err(R, t))
    return jac, err


def make_cases():
    cases = []
    rng = np.random.default_rng(2024)
    variants = [
        ("flags", {}, 0.03, 0.02),
        ("far_start", {}, 0.25, 0.15),                               # overshoots: rejected steps, damping grows
        ("tight_thresholds", dict(min_grad=1e-9, min_param_inc=1e-9, jac_thresh=1e-3), 0.05, 0.03),  # runs long, Jacobian-skip branch
        ("low_max_damp", dict(max_damp=1e-4, init_damp=1e-4, damp_inc=10.0), 0.3, 0.2),              # stops at max damping
        ("few_iters", dict(max_iters=3, min_grad=1e-9, min_param_inc=1e-9), 0.1, 0.05),              # stops at the iteration limit
        ("heavy_damp", dict(init_damp=1e-2, min_damp=1e-3, max_damp=1.0, damp_dec=2.0, damp_inc=5.0), 0.1, 0.1),
    ]
    for dof in (6, 7):
        for name, over, rot_mag, trans_mag in variants:
            n = 60
            p = np.stack([rng.uniform(-1.0, 1.0, n), rng.uniform(-0.8, 0.8, n), rng.uniform(1.5, 4.0, n)], 1)
            R_true = rodrigues(rng.standard_normal(3) * 0.05)
            t_true = rng.standard_normal(3) * 0.05
            s_true = 1.0 if dof == 6 else 1.15
            x = s_true * p @ R_true.T + t_true
            uv = np.stack([CAM[0] * x[:, 0] / x[:, 2] + CAM[2], CAM[1] * x[:, 1] / x[:, 2] + CAM[3]], 1) + rng.standard_normal((n, 2)) * 0.3
            dz = x[:, 2] + rng.standard_normal(n) * 0.01
            R0 = (rodrigues(rng.standard_normal(3) * rot_mag) @ R_true).astype(np.float32)
            t0 = (t_true + rng.standard_normal(3) * trans_mag).astype(np.float32)
            s0 = np.float32(1.0)
            opt = dict(FLAGS)
            opt.update(over)
            cases.append(dict(name=f"dof{dof}_{name}", dof=dof, pts=np.concatenate([p, uv, dz[:, None]], 1).astype(np.float64), R0=R0, t0=t0, s0=s0, opt=opt))
    return cases


def run_reference_loop(exe, c):
    o = c["opt"]
    head = [c["dof"], len(c["pts"]), o["init_damp"], o["min_damp"], o["max_damp"], o["damp_inc"], o["damp_dec"], o["jac_thresh"], o["min_grad"],
            o["min_param_inc"], o["max_iters"], *CAM, WZ[c["dof"]]]
    text = " ".join(repr(float(v)) if not isinstance(v, int) else str(v) for v in head) + "\n"
    text += " ".join(repr(float(v)) for v in np.concatenate([c["R0"].reshape(-1), c["t0"].reshape(-1), [c["s0"]]])) + "\n"
    text += "\n".join(" ".join(repr(float(v)) for v in row) for row in c["pts"]) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    kinds, rows, final = [], [], None
    for line in out:
        tok = line.split()
        if tok[0] == "F":
            final = np.array([float(v) for v in tok[1:]])
        else:
            kinds.append(tok[0])
            rows.append([float(v) for v in tok[1:]])
    return "".join(kinds), np.array(rows), final


def update_depth_cases():
    rng = np.random.default_rng(77)
    cases = []
    for name, (H, W, C) in (("ud_24x32_c32", (24, 32, 32)), ("ud_48x64_c16", (48, 64, 16)), ("ud_40x30_c8", (40, 30, 8))):
        cases.append(dict(name=name, H=H, W=W, C=C, bias=rng.uniform(0.5, 3.0, H * W).astype(np.float32),
                          jac=(0.2 * rng.standard_normal((H * W, C))).astype(np.float32), code=(0.5 * rng.standard_normal(C)).astype(np.float32),
                          scale=np.float32(rng.uniform(0.7, 1.6))))
    return cases


def run_reference_update_depth(exe, c):
    text = f"0 {c['H']} {c['W']} {c['C']} {float(c['scale'])!r}\n"
    for a in (c["bias"], c["jac"].reshape(-1), c["code"]):
        text += " ".join(repr(float(v)) for v in a) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
    return np.array([float(v) for v in out], np.float32)


def main():
    import build_loop_ref

    assert build_loop_ref.available(), "needs /root/reference"
    exe = build_loop_ref.build()
    out = {}
    for c in make_cases():
        kinds, rows, final = run_reference_loop(exe, c)
        n = c["name"]
        out[n + "/pts"], out[n + "/R0"], out[n + "/t0"], out[n + "/s0"] = c["pts"], c["R0"], c["t0"], c["s0"]
        out[n + "/opt"] = np.array([c["opt"][k] for k in sorted(c["opt"])], np.float64)
        out[n + "/kinds"], out[n + "/log"], out[n + "/final"] = np.array(kinds), rows, final
        print(f"{n}: {kinds}  final error {final[-2]:.6g} after {int(final[-1])} iterations")
    for c in update_depth_cases():  # Mapper::UpdateMap's depth write-back, the reference's own UpdateDepth
        out[c["name"] + "/dpt_map"] = run_reference_update_depth(exe, c)
        print(f"{c['name']}: {len(out[c['name'] + '/dpt_map'])} depths")
    out["opt_keys"] = np.array(sorted(FLAGS))
    out["cam"] = np.array(CAM)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

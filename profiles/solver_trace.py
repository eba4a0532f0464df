"""Phase timeline of the block Cholesky (SAGE_BA_SOLVER_TRACE): which block column waits for what, and how long the
factorisation of one column takes.  The solver's cost depends on the graph (keyframes, links, code size) only, so a 32-keyframe
problem on small images reproduces the bench's elimination exactly.

    SAGE_BA_SOLVER_TRACE=1 python profiles/solver_trace.py [K] [natural]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SAGE_BA_SOLVER_TRACE", "1")
if len(sys.argv) > 2 and sys.argv[2] == "natural":
    os.environ["SAGE_BA_NATURAL_ORDER"] = "1"
import sage_slam_b200 as sage  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kfs = sage.synthetic.make_scene(num_kf=K, W=48, H=32, L=2, F=32, C=32, back_connections=3, seed=1234, step=0.01, rot_step_deg=0.3)
pairs = sage.synthetic.ordered_pairs(kfs)
ctx = sage.Context(0)
dk = [sage.DeviceKeyframe(ctx, k) for k in kfs]
ba = sage.LocalBA(ctx, dk)
geo_loss = float(0.03 * np.mean(kfs[0].dpt_map_bias.astype(np.float64) ** 2))
for i, j in pairs:
    ba.add_photometric(i, j, [10.0, 9.0])
for i, j in pairs:
    ba.add_geometric(i, j, geo_loss, 0.1)
for k in range(K):
    ba.add_code_prior(k, 1e-3)
    ba.add_scale_prior(k, 1.0, 1e-2)
ba.fix(0, pose=True, scale=True)
ba.set_state([k.pose_wk for k in kfs], np.stack([k.code for k in kfs]), [k.dpt_scale for k in kfs], 1e-4)
ba.linearize()
ba.assemble()
for _ in range(3):
    ba.solve(1e-2)  # timing only: tiny images need not give a well-posed system
tr = np.zeros((K, 8), np.int64)
order = (C.c_int * K)()
ctx.check(ctx.lib.sage_ba_problem_solver_trace(ba.h, tr.ctypes.data_as(C.c_void_p), order))
t0 = tr[:, 0].min()
print("solver:", ba.solver_info())
print("pos kf  deps | start   loaded  deps_done factored solved  published | wait_us  update_us potrf_us trsm_us publish_us")
for q in range(K):
    s = (tr[q, :6] - t0) / 1e3
    wait = tr[q, 6] / 1e3
    print(f"{q:3d} {order[q]:3d} {tr[q,7]:4d} | {s[0]:7.1f} {s[1]:7.1f} {s[2]:8.1f} {s[3]:8.1f} {s[4]:7.1f} {s[5]:8.1f} | {wait:7.1f} {s[2]-s[1]-wait:9.1f} "
          f"{s[3]-s[2]:8.1f} {s[4]-s[3]:7.1f} {s[5]-s[4]:9.1f}")
print("total factorisation:", (tr[:, 5].max() - t0) / 1e3, "us")

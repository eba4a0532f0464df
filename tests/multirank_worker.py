"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun), NCCL.  Every rank builds the same small problem,
uploads only the keyframes its own factors touch, linearises its shard, exchanges through the library's own NCCL
communicator and runs three LM iterations; rank 0 writes what it saw."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(rank, world, local, out, deterministic=True):
    import torch

    import helpers
    import problem_case as pc
    import sage_slam_b200 as sage
    from sage_slam_b200 import local_ba

    kfs, pairs, factors = pc.build(6)
    K = len(kfs)
    stream = torch.cuda.Stream(device=local)
    with torch.cuda.stream(stream):
        ctx = sage.Context(local, stream=stream.cuda_stream)
        need = local_ba.needed_keyframes(pairs, rank, world)
        dk = [sage.DeviceKeyframe(ctx, k) if i in need else None for i, k in enumerate(kfs)]
        ba = sage.LocalBA(ctx, dk, rank=rank, world=world)
        ba.deterministic(deterministic)
        for i, j in pairs:
            ba.add_photometric(i, j, helpers.PHOTO_WEIGHTS[:pc.PRM["L"]])
        for i, j in pairs:
            ba.add_geometric(i, j, pc.geo_loss(kfs), 0.1)
        for i, j in pairs:
            loc, homo, uv = pc.matches(kfs, i, j)
            ba.add_reprojection(i, j, loc, homo, uv, 0.03 * pc.PRM["W"] ** 2, 0.1)
        for k in range(K):
            ba.add_code_prior(k, pc.CODE_W)
            ba.add_scale_prior(k, 1.0, pc.SCALE_W)
        ba.fix(0, pose=True, scale=True)
        ba.set_state([k.pose_wk for k in kfs], np.stack([k.code for k in kfs]), [k.dpt_scale for k in kfs], helpers.EPS)
        if world > 1:
            ba.enable_nccl()
        ba.linearize()  # + exchange
        buf = ba.factor_buffer().copy()
        offs, _, owners = ba.factor_offsets()
        cost0 = ba.assemble()
        delta = ba.solve(1e-4, want_delta=True)
        trace, damp = [], 1e-4
        for _ in range(3):
            c, cand, acc, damp = ba.lm_step(damp)
            trace.append((c, cand, float(acc)))
        poses, codes, scales = ba.get_state()
        torch.cuda.synchronize()
    if rank == 0:
        # factor outputs in order of addition, independent of the segment layout
        _, dims, _ = local_ba.factor_layout([f[0] for f in factors], pc.PRM["C"])
        flat = np.concatenate([buf[o:o + D * D + D + 2] for o, D in zip(offs, dims)])
        np.savez(out, flat=flat, cost0=cost0, delta=delta, trace=np.array(trace), poses=np.stack([np.concatenate([R.reshape(-1), t]) for R, t in poses]),
                 codes=codes, scales=scales, resident=len(need), owners=np.array(owners))


if __name__ == "__main__":
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    run(rank, world, local, sys.argv[1], deterministic=(len(sys.argv) < 3 or sys.argv[2] != "fast"))
    dist.barrier()
    dist.destroy_process_group()

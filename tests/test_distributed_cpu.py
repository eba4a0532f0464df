"""CPU, world_size 2, gloo: the multi-GPU host logic -- round-robin factor sharding, the packed factor buffer,
the sum all-reduce and the redundant dense assembly -- with the CPU oracle standing in for the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)
import problem_case as pc
from sage_slam_b200 import local_ba


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kfs, pairs, factors = pc.build(3)
    owned = local_ba.shard_factors(len(factors), rank, world)
    buf = torch.from_numpy(pc.oracle_buffer(kfs, factors, owned=set(owned)))
    local_ba.allreduce_sum(buf)
    H, g, cost = local_ba.assemble_dense(buf.numpy(), factors, len(kfs), pc.PRM["C"])
    np.savez(os.path.join(out, f"rank{rank}.npz"), buf=buf.numpy(), H=H, g=g, cost=cost, owned=np.array(owned))
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    kfs, pairs, factors = pc.build(3)
    ref = pc.oracle_buffer(kfs, factors)
    Href, gref, cref = local_ba.assemble_dense(ref, factors, len(kfs), pc.PRM["C"])
    r = [dict(np.load(os.path.join(tmp_path, f"rank{k}.npz"))) for k in range(world)]
    assert sorted(list(r[0]["owned"]) + list(r[1]["owned"])) == list(range(len(factors)))
    for k in range(world):
        np.testing.assert_array_equal(r[k]["buf"], ref)  # x + 0 == x: bit-identical to the single-rank buffer
        np.testing.assert_array_equal(r[k]["H"], Href)
        np.testing.assert_array_equal(r[k]["g"], gref)
        assert float(r[k]["cost"]) == cref


def test_sharding_is_a_partition_for_every_world_size():
    """f % world == rank: every factor has exactly one owner, loads differ by at most one, for 1/2/3/4/8 ranks and the bench's
    540 factors (180 pairs x 3 kinds) as well as awkward counts."""
    for n in (540, 1, 7, 23):
        for world in (1, 2, 3, 4, 8):
            owned = [local_ba.shard_factors(n, r, world) for r in range(world)]
            assert sorted(f for o in owned for f in o) == list(range(n))
            sizes = [len(o) for o in owned]
            assert max(sizes) - min(sizes) <= 1
    # the packed layout the all-reduce carries: one [AtA | Atb | error | inliers] block per factor, geometric blocks wider
    offs, dims, total = local_ba.factor_layout(["photo", "geo", "reproj"], 32)
    assert dims == [45, 78, 45] and offs == [0, 45 * 45 + 45 + 2, 45 * 45 + 45 + 2 + 78 * 78 + 78 + 2]
    assert total == 2 * (45 * 45 + 47) + 78 * 78 + 80


def _worker4(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kfs, pairs, factors = pc.build(3)
    owned = set(local_ba.shard_factors(len(factors), rank, world))
    buf = torch.from_numpy(pc.oracle_buffer(kfs, factors, owned=owned))
    local_ba.allreduce_sum(buf)
    if rank == 0:
        np.save(os.path.join(out, "buf4.npy"), buf.numpy())
    dist.destroy_process_group()


def test_four_rank_allreduce_reproduces_single_rank(tmp_path):
    mp.spawn(_worker4, args=(4, _free_port(), str(tmp_path)), nprocs=4, join=True)
    kfs, pairs, factors = pc.build(3)
    np.testing.assert_array_equal(np.load(os.path.join(tmp_path, "buf4.npy")), pc.oracle_buffer(kfs, factors))

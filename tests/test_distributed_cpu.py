"""CPU, world_size 2 and 4, gloo: the multi-GPU host logic -- keyframe-owner sharding, the segmented packed factor buffer,
the exchange and the redundant assembly -- with the CPU oracle standing in for the kernels."""
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
    kfs, pairs, factors = pc.build(4)
    K = len(kfs)
    owned = local_ba.shard_factors(factors, rank, world)
    need = local_ba.needed_keyframes(pairs, rank, world)
    # a rank only ever reads the keyframes its own pairs touch
    for f in owned:
        assert factors[f][1] in need and factors[f][2] in need
    buf = torch.from_numpy(pc.oracle_buffer(kfs, factors, owned=set(owned), world=world))
    local_ba.allreduce_sum(buf)  # every slot is non-zero on exactly one rank: the sum is the gather
    H, g, cost = local_ba.assemble_dense(buf.numpy(), factors, K, pc.PRM["C"], world=world)
    np.savez(os.path.join(out, f"rank{rank}.npz"), buf=buf.numpy(), H=H, g=g, cost=cost, owned=np.array(owned), need=np.array(sorted(need)))
    dist.destroy_process_group()


def _check(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    kfs, pairs, factors = pc.build(4)
    K = len(kfs)
    ref1 = pc.oracle_buffer(kfs, factors)  # single-rank layout
    Href, gref, cref = local_ba.assemble_dense(ref1, factors, K, pc.PRM["C"])
    refw = pc.oracle_buffer(kfs, factors, world=world)  # the same factors in the `world`-segment layout
    r = [dict(np.load(os.path.join(tmp_path, f"rank{k}.npz"))) for k in range(world)]
    assert sorted(f for x in r for f in x["owned"]) == list(range(len(factors)))
    for k in range(world):
        np.testing.assert_array_equal(r[k]["buf"], refw)  # x + 0 == x: bit-identical to one rank computing everything
        np.testing.assert_array_equal(r[k]["H"], Href)  # and the assembled system does not depend on the world size
        np.testing.assert_array_equal(r[k]["g"], gref)
        assert float(r[k]["cost"]) == cref


def test_two_rank_exchange_reproduces_single_rank(tmp_path):
    _check(2, tmp_path)


def test_four_rank_exchange_reproduces_single_rank(tmp_path):
    _check(4, tmp_path)


def test_sharding_is_a_partition_for_every_world_size():
    """Pairs sorted by (host, target) cut into equal runs: every factor has exactly one owner, all kinds of a pair share it, loads
    differ by at most one pair, a rank needs only a band of keyframes -- for the bench graph (32 keyframes, 3 back-connections,
    180 ordered pairs x 3 kinds) and awkward sizes, and the C library applies the same rule."""
    import ctypes as C

    import sage_slam_b200 as sage

    lib = sage.capi.load()
    for K, back in ((32, 3), (5, 2), (7, 6), (3, 1)):
        pairs = [(k, j) for k in range(K) for j in range(max(0, k - back), k)]
        pairs = [p for (a, b) in pairs for p in ((a, b), (b, a))]
        factors = [(kind, i, j) for kind in ("photo", "geo", "reproj") for (i, j) in pairs]
        for world in (1, 2, 3, 4, 8):
            owned = [local_ba.shard_factors(factors, r, world) for r in range(world)]
            assert sorted(f for o in owned for f in o) == list(range(len(factors)))
            own = local_ba.shard_owners(pairs, world)
            loads = [sum(1 for r in own.values() if r == q) for q in range(world)]
            assert max(loads) - min(loads) <= 1, (K, world, loads)
            for f, (_, i, j) in enumerate(factors):  # all kinds of a pair on one rank
                assert f in owned[own[(i, j)]]
            for r in range(world):
                need = local_ba.needed_keyframes(pairs, r, world)
                hosts = sorted({i for (i, j), q in own.items() if q == r})
                if hosts:
                    assert need <= set(range(max(0, hosts[0] - back), min(K, hosts[-1] + back + 1)))
            pi = (C.c_int * len(pairs))(*[p[0] for p in pairs])
            pj = (C.c_int * len(pairs))(*[p[1] for p in pairs])
            out = (C.c_int * len(pairs))()
            assert lib.sage_ba_shard_plan(len(pairs), pi, pj, world, out) == 0
            assert list(out) == [own[p] for p in pairs]
    # K = 32 on 8 ranks: 22 or 23 pairs each (180 / 8), at most 11 keyframes resident per rank instead of 32
    pairs = [p for k in range(32) for j in range(max(0, k - 3), k) for p in ((k, j), (j, k))]
    own = local_ba.shard_owners(pairs, 8)
    assert sorted(set(sum(1 for r in own.values() if r == q) for q in range(8))) == [22, 23]
    assert max(len(local_ba.needed_keyframes(pairs, r, 8)) for r in range(8)) <= 11
    # the packed layout: one [AtA | Atb | error | inliers] block per factor, geometric blocks wider; world > 1: equal segments
    offs, dims, total = local_ba.factor_layout(["photo", "geo", "reproj"], 32)
    assert dims == [45, 78, 45] and offs == [0, 45 * 45 + 45 + 2, 45 * 45 + 45 + 2 + 78 * 78 + 78 + 2]
    assert total == 2 * (45 * 45 + 47) + 78 * 78 + 80
    offs, dims, total = local_ba.factor_layout(["photo", "geo", "reproj"], 32, owners=[1, 0, 1], world=2)
    seg = (max(2 * (45 * 45 + 47), 78 * 78 + 80) + 31) // 32 * 32
    assert offs == [seg, 0, seg + 45 * 45 + 47] and total == 2 * seg

"""GPU, >= 2 devices: the sharded problem on real GPUs through the library's own NCCL collective (torchrun, one process per
GPU) against the same problem on one GPU.  Keyframe-owner sharding never splits a factor and a factor's CTA decomposition does
not depend on the rank count, so the exchanged factor buffer, the assembled system, the step and the whole LM trajectory must
be BIT-identical for every world size -- far inside BASELINE's gates (cost <= 1e-4 relative, pose update <= 1e-5)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(world, out, mode="deterministic"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(helpers.ROOT, "tests", "multirank_worker.py"), out, mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return dict(np.load(out))


@pytest.mark.gpu
def test_multi_gpu_lm_is_bit_identical_to_one_gpu(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import multirank_worker

    ref_out = str(tmp_path / "w1.npz")
    multirank_worker.run(0, 1, 0, ref_out)
    ref = dict(np.load(ref_out))
    for world in [w for w in (2, 4, 8) if w <= ngpu]:
        got = _launch(world, str(tmp_path / f"w{world}.npz"))
        assert got["resident"] <= 6  # rank 0 holds only the keyframes its pairs touch (all 6 on 2 ranks, 4 of 6 on 6+)
        assert len(set(got["owners"].tolist())) == world  # 24 ordered pairs: every rank owns some
        np.testing.assert_array_equal(got["flat"], ref["flat"])    # every factor's [AtA | Atb | error | inliers]
        assert float(got["cost0"]) == float(ref["cost0"])
        # BASELINE's gates first (what matters), then the stronger statement
        assert np.abs(got["delta"][:36] - ref["delta"][:36]).max() <= 1e-5
        assert np.abs(got["trace"][:, :2] - ref["trace"][:, :2]).max() <= 1e-4 * np.abs(ref["trace"][:, :2]).max()
        np.testing.assert_array_equal(got["delta"], ref["delta"])
        np.testing.assert_array_equal(got["trace"], ref["trace"])
        np.testing.assert_array_equal(got["poses"], ref["poses"])
        np.testing.assert_array_equal(got["codes"], ref["codes"])
        np.testing.assert_array_equal(got["scales"], ref["scales"])


@pytest.mark.gpu
def test_multi_gpu_default_slicing_stays_inside_the_gates(tmp_path):
    """Default (fastest) CTA decomposition: slices are sized per rank, so a factor's partial sums are added in a different order
    on 2 GPUs than on 1.  The results must still agree far inside BASELINE's gates: pose update <= 1e-5, cost <= 1e-4."""
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import multirank_worker

    ref_out = str(tmp_path / "f1.npz")
    multirank_worker.run(0, 1, 0, ref_out, deterministic=False)
    ref = dict(np.load(ref_out))
    for world in [w for w in (2, 4, 8) if w <= ngpu]:
        got = _launch(world, str(tmp_path / f"f{world}.npz"), "fast")
        scale = np.abs(ref["flat"]).max()
        assert np.abs(got["flat"] - ref["flat"]).max() <= 2e-6 * scale
        assert abs(float(got["cost0"]) - float(ref["cost0"])) <= 1e-6 * float(ref["cost0"])
        assert np.abs(got["delta"][:36] - ref["delta"][:36]).max() <= 1e-5
        assert np.abs(got["trace"][:, :2] - ref["trace"][:, :2]).max() <= 1e-4 * np.abs(ref["trace"][:, :2]).max()
        assert np.abs(got["poses"] - ref["poses"]).max() <= 1e-5

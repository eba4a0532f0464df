"""CPU: the input producers pinned by the reference's OWN functions.

tests/golden/producer_pins.npz = outputs of GenerateMaskPyramid (core/mapping/mapping_utils.cpp:321-342),
Mapper::GenerateGaussianPyramidWithGrad (core/mapping/mapper.cpp:1383-1426, with ComputeSpatialGrad and the constructor's Gaussian
kernel) and GenerateValidLocations (core/mapping/mapping_utils.h:258-296), extracted verbatim at build time and run with libtorch
on the CPU (oracle/build_loop_ref.py, producer_pins.cpp, make_golden_producers.py).  Held to them here: the host builders of
sage-slam_b200/frames.py (what the synthetic scenes, the tests and the bench feed the library) and the oracle's torch restatements;
the device builder (csrc/prep.cu, `feat_map` in sage_ba_keyframe_desc) is held to the host builder on the GPU
(tests/test_gpu_parity.py)."""
import os

import numpy as np
import pytest
import torch

import helpers
import make_golden_producers as G
import oracle as O
import sage_slam_b200 as sage

PINS = os.path.join(helpers.ROOT, "tests", "golden", "producer_pins.npz")
CASES = G.make_cases()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_host_builders_match_the_references_own_producers(case):
    pins = np.load(PINS)
    n, L = case["name"], case["L"]
    masks = sage.frames.mask_pyramid(case["mask"], L)
    np.testing.assert_array_equal(np.concatenate([m.reshape(-1) for m in masks]), pins[n + "/masks"])  # nearest resize: exact
    pyr, grad = sage.frames.gaussian_pyramid_with_grad(case["feat"], masks)
    F = case["F"]
    want_pyr, want_grad = pins[n + "/pyr"].reshape(F, -1), pins[n + "/grad"].reshape(2, F, -1)
    assert pyr.shape == want_pyr.shape and grad.shape == want_grad.shape
    # level 0 is the input itself and its central differences: exact; deeper levels go through one fp32 3x3 convolution per level
    hw = case["H"] * case["W"]
    np.testing.assert_array_equal(pyr[:, :hw], want_pyr[:, :hw])
    np.testing.assert_array_equal(grad[:, :, :hw], want_grad[:, :, :hw])
    np.testing.assert_allclose(pyr, want_pyr, rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(grad, want_grad, rtol=2e-6, atol=2e-7)
    loc, homo = sage.frames.valid_locations(case["mask"], case["cam"])
    np.testing.assert_array_equal(loc, pins[n + "/loc1d"])
    np.testing.assert_allclose(homo, pins[n + "/homo"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_oracle_restatements_match_the_references_own_producers(case):
    pins = np.load(PINS)
    n, L, F = case["name"], case["L"], case["F"]
    tm = O.mask_pyramid(torch.from_numpy(case["mask"])[None, None], L)
    np.testing.assert_array_equal(np.concatenate([m.reshape(-1).numpy() for m in tm]), pins[n + "/masks"])
    tp, tg = O.gaussian_pyramid_with_grad(torch.from_numpy(case["feat"])[None], tm)
    np.testing.assert_allclose(tp.numpy(), pins[n + "/pyr"].reshape(F, -1), rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(tg.numpy(), pins[n + "/grad"].reshape(2, F, -1), rtol=2e-6, atol=2e-7)
    loc, homo = O.valid_locations(case["mask"], case["cam"])
    np.testing.assert_array_equal(loc, pins[n + "/loc1d"])
    np.testing.assert_allclose(homo, pins[n + "/homo"], rtol=0, atol=1e-7)

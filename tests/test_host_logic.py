"""CPU: host-side pieces of the path -- camera pyramid, Gaussian pyramid + gradients, retraction, NearestPsd,
the tracker LM restatement, factor packing / dense assembly -- against independent restatements."""
import os

import numpy as np
import torch

import helpers
import oracle as O
import sage_slam_b200 as sage
from sage_slam_b200 import local_ba


def test_camera_pyramid_matches_oracle_and_halves():
    cam = [256.0, 250.0, 159.5, 127.5, 320, 256]
    a, b = sage.frames.camera_pyramid(cam, 4), O.camera_pyramid(cam, 4)
    np.testing.assert_array_equal(a, b)
    assert list(a[:, 4]) == [320, 160, 80, 40] and list(a[:, 5]) == [256, 128, 64, 32]
    np.testing.assert_allclose(a[1, :4], [128, 125, 79.75, 63.75])


def test_gaussian_pyramid_numpy_vs_torch_restatement():
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((5, 48, 64)).astype(np.float32)
    yy, xx = np.mgrid[0:48, 0:64]
    mask = ((xx - 32) ** 2 / 900 + (yy - 24) ** 2 / 500 <= 1).astype(np.float32)
    masks = sage.frames.mask_pyramid(mask, 4)
    pyr, grad = sage.frames.gaussian_pyramid_with_grad(feat, masks)
    tm = O.mask_pyramid(torch.from_numpy(mask)[None, None], 4)
    for a, b in zip(masks, tm):
        np.testing.assert_array_equal(a, b[0, 0].numpy())
    tp, tg = O.gaussian_pyramid_with_grad(torch.from_numpy(feat)[None], tm)
    np.testing.assert_allclose(pyr, tp.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(grad, tg.numpy(), rtol=1e-5, atol=1e-6)
    assert pyr.shape == (5, 64 * 48 + 32 * 24 + 16 * 12 + 8 * 6)


def test_valid_locations():
    cam = np.array([50.0, 50.0, 31.5, 23.5, 64, 48], np.float32)
    m = np.zeros((48, 64), np.float32)
    m[10, 20] = 1
    loc, homo = sage.frames.valid_locations(m, cam)
    assert list(loc) == [10 * 64 + 20]
    np.testing.assert_allclose(homo[0], [(20 - 31.5) / 50, (10 - 23.5) / 50, 1.0])
    l2, h2 = O.valid_locations(m, cam)
    np.testing.assert_array_equal(loc, l2)
    np.testing.assert_array_equal(homo, h2)


def test_se3_exp_is_left_multiplicative_and_orthonormal():
    w, v = np.array([0.02, -0.01, 0.03]), np.array([0.1, 0.2, -0.05])
    R, t = O.se3_exp(w, v)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    from scipy.linalg import expm

    X = np.zeros((4, 4))
    X[:3, :3] = O.so3_hat(w)
    X[:3, 3] = v
    E = expm(X)
    np.testing.assert_allclose(R, E[:3, :3], atol=1e-12)
    np.testing.assert_allclose(t, E[:3, 3], atol=1e-12)
    R0, t0 = O.se3_exp(np.array([0.3, 0.1, -0.2]), np.array([1.0, 2.0, 3.0]))
    R1, t1 = O.retract(R0, t0, np.concatenate([v, w]))
    np.testing.assert_allclose(R1, R @ R0, atol=1e-12)
    np.testing.assert_allclose(t1, R @ t0 + t, atol=1e-12)


def test_nearest_psd_reference_quirk():
    """the reference forms V^T S V, which is NOT the identity map on PSD input."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((8, 8))
    B = A @ A.T
    np.testing.assert_allclose(O.nearest_psd(B, reference_faithful=False), B, rtol=1e-9, atol=1e-9)
    Q = O.nearest_psd(B, reference_faithful=True)
    assert np.linalg.norm(Q - B) / np.linalg.norm(B) > 1e-2
    assert np.linalg.eigvalsh(Q).min() > -1e-9


def test_tracker_lm_restatement_converges_on_a_quadratic():
    """camera_tracker.cpp:1156-1279 bookkeeping on a synthetic 6-DoF quadratic bowl."""
    rng = np.random.default_rng(5)
    J = rng.standard_normal((40, 6))
    target = np.array([0.02, -0.01, 0.03, 0.01, -0.02, 0.015])

    def params(R, t):
        return np.concatenate([t, O.rotation_to_angle_axis(R)])

    def jac(R, t):
        r = J @ (target - params(R, t))
        return J.T @ J, J.T @ r, float(r @ r)

    def err(R, t):
        r = J @ (target - params(R, t))
        return float(r @ r)

    R, t, e, trace = O.tracker_lm(jac, err, np.eye(3, dtype=np.float32), np.zeros(3, np.float32), max_iters=30)
    assert e < 1e-6 * err(np.eye(3), np.zeros(3)) and all(a for _, _, a, _ in trace[:-1])


def test_factor_packing_and_dense_assembly_layout():
    K, C = 3, 8
    factors = [("photo", 0, 1), ("geo", 1, 2), ("reproj", 2, 0)]
    offs, dims, total = local_ba.factor_layout([f[0] for f in factors], C)
    assert dims == [21, 30, 21] and offs == [0, 21 * 21 + 23, 21 * 21 + 23 + 30 * 30 + 32]
    buf = np.zeros(total, np.float32)
    rng = np.random.default_rng(1)
    mats = []
    for (kind, i, j), off, D in zip(factors, offs, dims):
        A = rng.standard_normal((D, D)).astype(np.float32)
        A = A @ A.T
        b = rng.standard_normal(D).astype(np.float32)
        local_ba.pack_factor(buf, off, D, A, b, 1.5, 10)
        mats.append((A, b))
    H, g, cost = local_ba.assemble_dense(buf, factors, K, C)
    assert cost == 4.5 and H.shape == (K * (7 + C),) * 2
    np.testing.assert_allclose(H, H.T)
    # photometric 0->1: pose0 block lands on pose 0, scale column on the scale slot of keyframe 0
    np.testing.assert_allclose(H[0:6, 0:6], mats[0][0][0:6, 0:6] + mats[2][0][6:12, 6:12], rtol=1e-6)
    s0 = 6 * K + 0 * (C + 1) + C
    np.testing.assert_allclose(g[s0], mats[0][1][12 + C], rtol=1e-6)
    # geometric 1->2 couples code_1 with code_2
    c1, c2 = 6 * K + 1 * (C + 1), 6 * K + 2 * (C + 1)
    np.testing.assert_allclose(H[c1:c1 + C, c2:c2 + C], mats[1][0][12:12 + C, 12 + C:12 + 2 * C], rtol=1e-6)
    # pair sharding: three distinct pairs on 3 ranks, one each (sorted by host keyframe)
    assert [local_ba.shard_factors(factors, r, 3) for r in range(3)] == [[0], [1], [2]]


def test_factor_partition_and_nearest_psd_match_oracle():
    """factors._partition == the reference's HessianFactor slicing; factors.nearest_psd == the oracle restatement."""
    from sage_slam_b200 import factors

    rng = np.random.default_rng(9)
    C = 8
    D = 13 + C
    A = rng.standard_normal((D, D))
    A = (A @ A.T).astype(np.float32)
    b = rng.standard_normal(D).astype(np.float32)
    for mode, faithful in (("reference", True), ("exact", False)):
        np.testing.assert_allclose(factors.nearest_psd(A, mode), O.nearest_psd(A, reference_faithful=faithful), rtol=1e-12, atol=1e-12)
    keys = [factors.pose_key(3), factors.pose_key(5), factors.code_key(3), factors.scale_key(3)]
    hf = factors._partition(A, b, keys, [6, 6, C, 1], 2.5, "none")
    assert len(hf.Gs) == 10 and len(hf.gs) == 4 and hf.f == 2.5
    assert [g.shape for g in hf.Gs] == [(6, 6), (6, 6), (6, C), (6, 1), (6, 6), (6, C), (6, 1), (C, C), (C, 1), (1, 1)]
    np.testing.assert_array_equal(hf.Gs[2], A.astype(np.float64)[0:6, 12:12 + C])       # G13
    np.testing.assert_array_equal(hf.Gs[8], A.astype(np.float64)[12:12 + C, 12 + C:])   # G34
    G, g = hf.information()
    np.testing.assert_allclose(G, (A.astype(np.float64) + A.astype(np.float64).T) / 2, rtol=1e-6)
    np.testing.assert_array_equal(g, b.astype(np.float64))
    hg = factors._partition(np.eye(14 + 2 * C), np.zeros(14 + 2 * C), list(range(6)), [6, 6, C, C, 1, 1], 0.0, "none")
    assert len(hg.Gs) == 21


def test_bench_accounting_matches_the_survey():
    """bench.py's roofline numerator is SURVEY.md section 8(d)'s per-pair figure: 68.2 / 40.3 / 23.9 MB at 320x256, F = C = 32,
    L = 4, dense sampling; the clock sampler degrades to a well-formed record where neither NVML nor nvidia-smi has a GPU."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("sage_bench", os.path.join(helpers.ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.CONFIGS[3]
    photo, photo_err, geo = bench.algorithmic_bytes(wl)
    assert abs(photo / 1e6 - 68.2) < 0.1 and abs(photo_err / 1e6 - 40.3) < 0.1 and abs(geo / 1e6 - 23.9) < 0.1
    assert (wl["num_kf"], wl["W"], wl["H"], wl["F"], wl["C"], wl["L"]) == (32, 320, 256, 32, 32, 4)  # BASELINE configs[3]
    # the other BASELINE configurations bench.py can run: configs[0] (2 KF, 128x96, F16, C8, photometric only: 4.87 MB per pair),
    # configs[2] (16 KF, full covisibility), configs[4] (256 KF, average degree 8)
    c0, c2, c4 = bench.CONFIGS[0], bench.CONFIGS[2], bench.CONFIGS[4]
    assert (c0["num_kf"], c0["W"], c0["H"], c0["F"], c0["C"], c0["kinds"]) == (2, 128, 96, 16, 8, ("photo",))
    assert abs(bench.algorithmic_bytes(c0)[0] / 1e6 - 4.87) < 0.02
    assert (c2["num_kf"], c2["graph"]) == (16, "full") and (c4["num_kf"], c4["graph"]) == (256, "sparse8")
    s = bench.ClockSampler(0)
    s.start()
    s.mark_begin()
    s.mark_end()
    c = s.stop()
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}


def test_truncation_tf32_split_error_bound():
    """The kernels' 3xTF32 rank-k update splits x = hi + lo by truncation (sage_common.cuh split_tf32).  Restated in numpy:
    hi + lo reproduces x to 2^-20, and hi*hi' + hi*lo' + lo*hi' (what the three MMAs add up) reproduces the fp32 product sum to
    ~1e-6 relative -- two orders inside the 1e-4 parity gate."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-8, 8, 200000))).astype(np.float32)
    y = (rng.standard_normal(200000) * np.exp(rng.uniform(-8, 8, 200000))).astype(np.float32)

    def split(v):
        hi = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        lo = ((v - hi).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        return hi, lo

    xh, xl = split(x)
    yh, yl = split(y)
    assert np.all(np.abs(x.astype(np.float64) - xh - xl.astype(np.float64)) <= 2.0 ** -20 * np.abs(x))
    # TF32 operands have 10 mantissa bits: both parts must be exactly representable
    for part in (xh, xl):
        assert np.all((part.view(np.uint32) & np.uint32(0x1FFF)) == 0)
    prod = xh.astype(np.float64) * yh + xh.astype(np.float64) * yl + xl.astype(np.float64) * yh
    exact = x.astype(np.float64) * y
    assert np.all(np.abs(prod - exact) <= 4e-6 * np.abs(exact))
    # a Gram-matrix entry: positive terms, errors do not accumulate beyond the per-term bound
    assert abs((xh.astype(np.float64) ** 2 + 2 * xh.astype(np.float64) * xl).sum() / (x.astype(np.float64) ** 2).sum() - 1.0) < 2e-6


def test_std_shuffle_reproduces_libstdcxx():
    """frames.std_shuffle against std::shuffle + std::mt19937 of g++ 13 (tests/golden/std_shuffle_gcc13.txt, generated by
    tests/golden/make_std_shuffle.cpp): both code paths (two swaps per engine call for n <= 65535, one otherwise)."""
    from sage_slam_b200 import frames

    n_cases = 0
    for line in open(os.path.join(helpers.ROOT, "tests", "golden", "std_shuffle_gcc13.txt")):
        if line.startswith("#"):
            continue
        left, h = line.split("|")
        t = [int(v) for v in left.split()]
        n, seed, first = t[0], t[1], t[2:]
        idx = frames.std_shuffle(n, seed)
        assert list(idx[:len(first)]) == first
        hh = 1469598103934665603
        for v in idx.tolist():
            hh = ((hh ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert hh == int(h)
        n_cases += 1
    assert n_cases == 7
    old = frames.std_shuffle(1000, 42, libstdcxx="9")  # GCC <= 10 variant: a permutation, different draw
    assert sorted(old.tolist()) == list(range(1000)) and list(old[:8]) != list(frames.std_shuffle(1000, 42)[:8])


def test_avg_squared_dpt_bias_is_the_masked_mean():
    """Mapper::BuildKeyframe (core/mapping/mapper.cpp:1376-1378): sum((bias * mask)^2) / sum(mask), not the mean over all pixels --
    the two differ under an endoscope mask, and the value scales the robust loss of every geometric factor."""
    from sage_slam_b200 import mapper

    kfs = sage.synthetic.make_scene(num_kf=1, W=64, H=48, L=2, F=16, C=8, mask="ellipse", seed=3)
    kf = kfs[0]
    m = kf.video_mask.reshape(-1).astype(np.float64)
    b = kf.dpt_map_bias.astype(np.float64)
    want = np.sum((b * m) ** 2) / np.sum(m)
    got = mapper.avg_squared_dpt_bias(kf)
    assert abs(got - want) <= 1e-5 * want
    assert abs(np.mean(b ** 2) - want) > 1e-2 * want  # the plain mean is something else here


def test_tcgen05_operand_order_and_finalize_algebra():
    """csrc/geometric.cu, geo_tc_kernel: operand columns [loB | hi | loA], ONE product D = A^T B with A = [hi | loA] (128 columns) and
    B = [loB | hi] (112 columns); the finalize kernel rebuilds J^T J = hi^T hi + X + X^T (X = hi^T lo) from D alone.  Restated in
    numpy with the kernel's index formulas: the result must equal the 3xTF32 sum hi^T hi + hi^T lo + lo^T hi exactly (fp64)."""
    rng = np.random.default_rng(3)
    WP, L1 = 80, 48
    L2 = WP - L1
    NC = WP + L2
    rows = rng.standard_normal((256, WP)).astype(np.float32)
    hi = (rows.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = ((rows - hi).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    assert np.abs(rows - hi - lo).max() <= 2.0 ** -20 * np.abs(rows).max()  # what the split drops
    hi, lo = hi.astype(np.float64), lo.astype(np.float64)
    stage = np.concatenate([lo[:, L1:], hi, lo[:, :L1]], axis=1)  # [loB | hi | loA]: 160 operand columns per sample
    A, B = stage[:, L2:L2 + 128], stage[:, :NC]
    assert A.shape[1] == 128 and B.shape[1] == 112
    P = A.T @ B  # the accumulator image the CTA flushes: [128][112]

    def xt(a, b):  # X^T[a][b] = (lo^T hi)[a][b], as in geo_finalize_kernel<.., TCP = true>
        return P[WP + a, L2 + b] if a < L1 else P[b, a - L1]

    S = np.empty((WP, WP))
    for r in range(WP):
        for c in range(WP):
            lo_i, hi_i = min(r, c), max(r, c)
            S[r, c] = P[lo_i, L2 + hi_i] + (xt(lo_i, hi_i) + xt(hi_i, lo_i))
    want = hi.T @ hi + hi.T @ lo + lo.T @ hi
    np.testing.assert_allclose(S, want, rtol=1e-12, atol=1e-12)
    # and the staging address formula: 32 four-byte stores of one instruction (4 consecutive samples x 8 channel quads) hit 32 banks
    SBO, LBO = 272, 144
    banks = {(((2 + (gl >> 1)) * SBO + (gl & 1) * 64 + q * 4) // 4) % 32 for q in range(4) for gl in range(8)}
    assert len(banks) == 32 and (20 * SBO // 4) % 32 == 16 and LBO + 128 <= SBO

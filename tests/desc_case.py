"""Seeded descriptor-map pairs for the cycle-matching tests (row f3).  desc1 is desc0 seen through a small smooth
displacement plus noise, so most keypoints have a cycle-consistent match and a few do not."""
import numpy as np

# name -> (C, H, W, K, shift (dy, dx), noise, thresh)
CASES = {
    "d16_small": (16, 48, 64, 96, (2, 3), 0.02, 2.0),
    "d32_ragged": (32, 40, 56, 77, (1, -2), 0.60, 2.0),
    "d16_native": (16, 96, 128, 256, (-3, 4), 0.45, 2.0),
    "d8_loose": (8, 32, 48, 50, (0, 1), 0.30, 3.0),
}


def _smooth(rng, C, H, W, octaves=3):
    out = np.zeros((C, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for o in range(octaves):
        f = 2.0 ** o
        for c in range(C):
            a, b, p = rng.uniform(-1, 1, 3)
            out[c] += (np.sin(f * (a * xx / 7.0 + b * yy / 5.0) + 6.28 * p) / f).astype(np.float32)
    return out


def build(name):
    C, H, W, K, shift, noise, thresh = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    d0 = _smooth(rng, C, H, W) + 0.3 * rng.standard_normal((C, H, W)).astype(np.float32)
    d1 = np.roll(d0, shift, axis=(1, 2)) + noise * rng.standard_normal((C, H, W)).astype(np.float32)
    kp = np.sort(rng.permutation(H * W)[:K]).astype(np.int64)
    kp = rng.permutation(kp)  # the reference's keypoints come from a shuffle: unsorted
    return {"desc0": np.ascontiguousarray(d0, np.float32), "desc1": np.ascontiguousarray(d1, np.float32), "kp": kp,
            "C": C, "H": H, "W": W, "K": K, "thresh": thresh, "shift": shift}


def signature(c):
    return np.array([c["desc0"].astype(np.float64).sum(), c["desc1"].astype(np.float64).sum(), float(c["kp"].sum())])

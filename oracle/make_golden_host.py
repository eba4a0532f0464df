"""Generate tests/golden/host_pins.npz with the reference's OWN host code (oracle/_ref/host_pins, see build_host_ref.py):
NearestPsd (core/mapping/mapping_utils.h:104-128, fp64, as every GTSAM-path factor calls it), se3_exp (:316-346, double and
float) and the tracker's damped colPivHouseholderQr solve (core/system/camera_tracker.cpp:1182-1183 / :1523-1524, float).
Runs here (CPU only, needs /root/reference); inputs are stored with the outputs."""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_host_ref  # noqa: E402


def gram(rng, n, m, scales):
    J = rng.standard_normal((m, n)) * scales[None, :]
    return J.T @ J / m


def cases():
    rng = np.random.default_rng(2024)
    psd = []
    for n, C in ((21, 8), (29, 16), (30, 8), (45, 32), (46, 16), (78, 32)):
        # column scales like a factor's AtA: pose columns ~ f/z, code / scale columns orders of magnitude smaller
        scales = np.concatenate([np.full(12, 50.0), np.full(n - 12, 0.3)]) * np.exp(rng.uniform(-1, 1, n))
        psd.append(gram(rng, n, 4 * n, scales))                       # full rank
        psd.append(gram(rng, n, n // 2, scales))                      # rank deficient
    B = rng.standard_normal((21, 21))
    psd.append((B + B.T) / 2)                                          # indefinite: the shift loop runs
    psd.append(gram(rng, 21, 60, np.ones(21)) + 1e-3 * rng.standard_normal((21, 21)))  # not symmetric
    psd.append(np.diag(np.arange(1.0, 9.0)))                           # diagonal, distinct eigenvalues
    se3 = [np.zeros(6), np.array([1e-9, 0, 0, 0.1, 0.2, 0.3]), np.array([0, 0, 0, 1.0, -2.0, 0.5])]
    se3 += [np.concatenate([rng.standard_normal(3) * s, rng.standard_normal(3)]) for s in (1e-6, 1e-3, 0.05, 0.5, 2.0, 3.1)]
    qr = []
    for n in (6, 7):
        for kind in ("well", "ill", "deficient"):
            m = {"well": 40, "ill": 40, "deficient": n - 2}[kind]
            sc = np.ones(n) if kind != "ill" else np.logspace(0, -4, n)
            A = gram(rng, n, m, sc).astype(np.float32)
            b = (rng.standard_normal(n) * np.sqrt(np.diag(A) + 1e-12)).astype(np.float32)
            for damp in (1e-6, 1e-4, 1e-2, 1.0, 100.0):
                qr.append((A, b, np.float32(damp)))
    return psd, se3, qr


def main():
    exe = build_host_ref.build()
    psd, se3, qr = cases()
    blob = b""
    for M in psd:
        blob += b"P" + struct.pack("<i", M.shape[0]) + np.ascontiguousarray(M, np.float64).tobytes()
    for x in se3:
        blob += b"E" + np.asarray(x, np.float64).tobytes()
        blob += b"F" + np.asarray(x, np.float32).tobytes()
    for A, b, d in qr:
        blob += b"Q" + struct.pack("<i", A.shape[0]) + A.tobytes() + b.tobytes() + np.float32(d).tobytes()
    out = subprocess.run([exe], input=blob, capture_output=True, check=True).stdout
    res, pos = {}, 0
    for k, M in enumerate(psd):
        n = M.shape[0]
        res[f"psd_in_{k}"] = M
        res[f"psd_out_{k}"] = np.frombuffer(out, np.float64, n * n, pos).reshape(n, n).copy()
        pos += 8 * n * n
    for k, x in enumerate(se3):
        res[f"se3_in_{k}"] = np.asarray(x, np.float64)
        res[f"se3_out64_{k}"] = np.frombuffer(out, np.float64, 12, pos).copy()
        pos += 96
        res[f"se3_out32_{k}"] = np.frombuffer(out, np.float32, 12, pos).copy()
        pos += 48
    for k, (A, b, d) in enumerate(qr):
        n = A.shape[0]
        res[f"qr_A_{k}"], res[f"qr_b_{k}"], res[f"qr_damp_{k}"] = A, b, d
        res[f"qr_x_{k}"] = np.frombuffer(out, np.float32, n, pos).copy()
        pos += 4 * n
    assert pos == len(out), (pos, len(out))
    res["counts"] = np.array([len(psd), len(se3), len(qr)])
    fn = os.path.join(os.path.dirname(HERE), "tests", "golden", "host_pins.npz")
    np.savez_compressed(fn, **res)
    print(fn, len(psd), len(se3), len(qr), os.path.getsize(fn))


if __name__ == "__main__":
    main()

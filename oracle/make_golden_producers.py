"""Generate tests/golden/producer_pins.npz: outputs of the reference's OWN input producers.

    python oracle/make_golden_producers.py      (only where /root/reference exists; see oracle/build_loop_ref.py, producer_pins.cpp)

GenerateMaskPyramid, Mapper::GenerateGaussianPyramidWithGrad (+ ComputeSpatialGrad, the constructor's Gaussian kernel) and
GenerateValidLocations, extracted verbatim at build time and run with libtorch on the CPU, on seeded feature maps under a full and
an endoscope (ellipse) mask.  tests/test_producer_pins.py holds the host builders (sage-slam_b200/frames.py), the oracle's
restatements and -- on the GPU -- the device builder of csrc/prep.cu to these outputs.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "..", "tests", "golden", "producer_pins.npz")


def make_cases():
    rng = np.random.default_rng(5150)
    cases = []
    for name, (F, H, W, L), ellipse in (("full_4x48x64_l4", (4, 48, 64, 4), False), ("ellipse_3x40x56_l3", (3, 40, 56, 3), True),
                                        ("ellipse_2x64x80_l4", (2, 64, 80, 4), True)):
        yy, xx = np.mgrid[0:H, 0:W]
        mask = np.ones((H, W), np.float32) if not ellipse else \
            (((xx - W / 2 + 0.5) / (0.46 * W)) ** 2 + ((yy - H / 2 + 0.5) / (0.44 * H)) ** 2 <= 1.0).astype(np.float32)
        feat = rng.standard_normal((F, H, W)).astype(np.float32)
        cam = np.array([0.9 * W, 0.95 * W, W / 2 - 0.5, H / 2 - 0.5, W, H], np.float32)
        cases.append(dict(name=name, F=F, H=H, W=W, L=L, mask=mask, feat=feat, cam=cam))
    return cases


def run_reference(exe, c):
    text = f"{c['F']} {c['H']} {c['W']} {c['L']} " + " ".join(repr(float(v)) for v in c["cam"][:4]) + "\n"
    text += " ".join(repr(float(v)) for v in c["mask"].reshape(-1)) + "\n" + " ".join(repr(float(v)) for v in c["feat"].reshape(-1)) + "\n"
    tok = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
    out, i = {}, 0
    while i < len(tok):
        tag, n = tok[i], int(tok[i + 1])
        out[tag] = np.array([float(v) for v in tok[i + 2:i + 2 + n]])
        i += 2 + n
    return out


def main():
    import build_loop_ref

    assert build_loop_ref.available(), "needs /root/reference"
    exe = build_loop_ref.build_producers()
    out = {}
    for c in make_cases():
        r = run_reference(exe, c)
        n = c["name"]
        out[n + "/masks"], out[n + "/pyr"], out[n + "/grad"] = r["M"].astype(np.float32), r["P"].astype(np.float32), r["G"].astype(np.float32)
        out[n + "/loc1d"], out[n + "/homo"] = r["L"].astype(np.int64), r["H"].astype(np.float32).reshape(-1, 3)
        print(f"{n}: {len(r['M'])} mask values, pyramid {len(r['P'])}, gradient {len(r['G'])}, {len(r['L'])} valid locations")
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

"""Generate tests/golden/retract_pins.npz: the reference's OWN SE(3) manifold for the mapping variables.

    python oracle/make_golden_retract.py      (only where /root/reference exists; see oracle/build_host_ref.py, retract_pins.cpp)

gtsam::traits<Sophus::SE3<Scalar>>::Retract / Local (core/gtsam/gtsam_traits.h:45-89), the struct extracted verbatim at build time
and compiled against the Sophus + Eigen the reference vendors, on seeded poses and increments: small steps, LM-sized steps, large
rotations and the zero increment, in float and in double.  tests/test_host_pins.py holds oracle.retract (the checker of the batched
LM's state update) and the library's sage_ba_se3_exp composition to them.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "..", "tests", "golden", "retract_pins.npz")


def make_cases():
    from make_golden_loop import rodrigues

    rng = np.random.default_rng(909)
    rows = []
    for k in range(40):
        R = rodrigues(rng.standard_normal(3) * (0.5 if k % 4 else 2.5))
        t = rng.standard_normal(3) * (0.3 if k % 3 else 5.0)
        mag = (1e-6, 1e-3, 5e-2, 1.0)[k % 4]
        delta = np.concatenate([rng.standard_normal(3) * mag, rng.standard_normal(3) * mag])
        if k == 7:
            delta[:] = 0.0
        if k == 11:
            delta[3:] = 0.0  # pure translation increment
        rows.append(np.concatenate([R.reshape(-1), t, delta]))
    return np.array(rows)


def main():
    import build_host_ref

    assert build_host_ref.retract_available(), "needs /root/reference"
    exe = build_host_ref.build_retract()
    cases = make_cases()
    text = f"{len(cases)}\n" + "\n".join(" ".join(repr(float(v)) for v in r) for r in cases) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    vals = np.array([[float(v) for v in line.split()] for line in out])
    assert vals.shape == (len(cases), 12 * 3 + 6)
    np.savez_compressed(OUT, cases=cases, pose32=vals[:, :12], retract32=vals[:, 12:24], retract64=vals[:, 24:36], local32=vals[:, 36:42])
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;  max |Local(p, Retract(p, d)) - d| =",
          np.abs(vals[:, 36:42] - cases[:, 12:18]).max())


if __name__ == "__main__":
    main()

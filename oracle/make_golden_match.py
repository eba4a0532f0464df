"""Generate tests/golden/match_pins.npz: the reference's OWN keypoint draw + descriptor cycle-matching statements on the seeded
descriptor pairs of tests/desc_case.py (row f3).

    python oracle/make_golden_match.py      (only where /root/reference exists; see oracle/build_loop_ref.py, match_pins.cpp)

The statements of ReprojectionFactor's constructor (core/gtsam/reprojection_factor.cpp:42-89), extracted verbatim at build time and
run with libtorch on the CPU: std::shuffle + std::mt19937 seeded with kf id x frame id picks the keypoints among the valid
locations, then the two dense descriptor searches and the cycle-consistency test.  tests/test_descriptor.py holds
frames.std_shuffle, the torch replay behind tests/golden/desc_*.npz and oracle.cycle_match to these outputs.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests")]
import desc_case  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "match_pins.npz")
IDS = {"d16_small": (3, 7), "d32_ragged": (12, 5), "d16_native": (41, 40), "d8_loose": (2, 9)}  # keyframe id, frame id -> the seed


def valid_locations(c):
    """All pixels except a 2-pixel border: the population the keypoints are drawn from."""
    H, W = c["H"], c["W"]
    yy, xx = np.mgrid[2:H - 2, 2:W - 2]
    return (yy * W + xx).reshape(-1).astype(np.int64)


def run_reference(exe, c, name):
    loc = valid_locations(c)
    kf_id, fr_id = IDS[name]
    text = f"{c['C']} {c['H']} {c['W']} {c['K']} {kf_id} {fr_id} {c['thresh']!r} {len(loc)}\n"
    text += " ".join(str(int(v)) for v in loc) + "\n"
    for d in (c["desc0"], c["desc1"]):
        text += " ".join(repr(float(v)) for v in d.reshape(-1)) + "\n"
    tok = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
    out, i = {}, 0
    while i < len(tok):
        tag, n = tok[i], int(tok[i + 1])
        out[tag] = np.array([int(v) for v in tok[i + 2:i + 2 + n]], np.int64)
        i += 2 + n
    return out


def main():
    import build_loop_ref

    assert build_loop_ref.available(), "needs /root/reference"
    exe = build_loop_ref.build_match()
    out = {}
    for name in desc_case.CASES:
        c = desc_case.build(name)
        r = run_reference(exe, c, name)
        out[name + "/keypoint_indexes"], out[name + "/raw_matched_locations_1d_1"] = r["I"], r["R"]
        out[name + "/cyc_matched_locations_1d_0"], out[name + "/matched_keypoint_indexes"] = r["C"], r["M"]
        print(f"{name}: {len(r['I'])} keypoints drawn, {len(r['M'])} cycle-consistent")
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

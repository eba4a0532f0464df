"""Golden vectors for the descriptor cycle-matching.

The reference implements this step as a chain of torch calls inside its factor constructors
(core/gtsam/reprojection_factor.cpp:57-92; identical in match_geometry_factor.cpp:62-97 and
camera_tracker.cpp:798-834); df_core itself cannot be built here (GTSAM/OpenCV/TEASER), so this script
replays those torch calls one for one (same ops, same order, same dtypes) on the seeded inputs of
tests/desc_case.py and stores only the index outputs:

    python oracle/make_golden_desc.py [cuda|cpu]      -> tests/golden/desc_<device>.npz (or gpurun_out/golden/ on a GPU box)

Both the CPU run (made in the build container) and the CUDA run (made on a B200 with gpurun) are committed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import desc_case  # noqa: E402


def reference_chain(feature_desc_0, feature_desc_1, keypoint_locations_1d, width, height, cyc_consis_thresh):
    # line numbers: core/gtsam/reprojection_factor.cpp
    from torch import index_select
    num_keypoints_ = keypoint_locations_1d.size(0)
    keypoint_locations_2d_x = torch.fmod(keypoint_locations_1d, float(width))                              # :58
    keypoint_locations_2d_y = torch.floor(keypoint_locations_1d / float(width))                            # :59
    channel = feature_desc_0.size(1)                                                                       # :63
    keypoint_features_0 = feature_desc_0.reshape(channel, height * width)[:, keypoint_locations_1d]        # :66
    feature_response_1 = -torch.sum(torch.square(keypoint_features_0.reshape(channel, num_keypoints_, 1) -
                                                 feature_desc_1.reshape(channel, 1, height * width)), 0, False)   # :68-70
    raw_matched_locations_1d_1 = torch.max(feature_response_1, 1, False)[1]                                # :72
    raw_matched_features_1 = feature_desc_1.reshape(channel, height * width)[:, raw_matched_locations_1d_1]  # :74
    feature_response_0 = -torch.sum(torch.square(raw_matched_features_1.reshape(channel, num_keypoints_, 1) -
                                                 feature_desc_0.reshape(channel, 1, height * width)), 0, False)   # :77-79
    cyc_matched_locations_1d_0 = torch.max(feature_response_0, 1, False)[1]                                # :82
    cyc_matched_locations_2d_x = torch.fmod(cyc_matched_locations_1d_0, float(width))                      # :83
    cyc_matched_locations_2d_y = torch.floor(cyc_matched_locations_1d_0 / float(width))                    # :84
    cyc_distances_sq = torch.square(keypoint_locations_2d_x - cyc_matched_locations_2d_x) + \
        torch.square(keypoint_locations_2d_y - cyc_matched_locations_2d_y)                                 # :86-87
    inlier_keypoint_indexes = torch.nonzero(cyc_distances_sq <= (cyc_consis_thresh * cyc_consis_thresh)).reshape(-1)  # :89
    matched_locations_1d_1 = raw_matched_locations_1d_1[inlier_keypoint_indexes].to(torch.int32)           # :101
    matched_locations_2d_1 = torch.stack([torch.fmod(matched_locations_1d_1, float(width)),
                                          torch.floor(matched_locations_1d_1 / float(width))], 1)          # :104-105
    return {"raw_matched_locations_1d_1": raw_matched_locations_1d_1, "cyc_matched_locations_1d_0": cyc_matched_locations_1d_0,
            "inlier_within_keypoint_indexes": inlier_keypoint_indexes, "matched_locations_1d_1": matched_locations_1d_1,
            "matched_locations_2d_1": matched_locations_2d_1}


def main():
    devname = sys.argv[1] if len(sys.argv) > 1 else ("cuda" if torch.cuda.is_available() else "cpu")
    dev = torch.device(devname)
    out = {}
    with torch.no_grad():
        for name in desc_case.CASES:
            c = desc_case.build(name)
            d0 = torch.from_numpy(c["desc0"][None]).to(dev)
            d1 = torch.from_numpy(c["desc1"][None]).to(dev)
            kp = torch.from_numpy(c["kp"]).to(dev)
            r = reference_chain(d0, d1, kp, c["W"], c["H"], c["thresh"])
            for k, v in r.items():
                out[f"{name}/{k}"] = v.cpu().numpy()
            out[f"{name}/sig"] = desc_case.signature(c)
    outdir = os.path.join(ROOT, "gpurun_out", "golden") if devname == "cuda" else os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    fn = os.path.join(outdir, f"desc_{devname}.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, {k: v.shape for k, v in out.items() if k.endswith("inlier_within_keypoint_indexes")})


if __name__ == "__main__":
    main()

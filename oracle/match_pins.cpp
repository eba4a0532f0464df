// match_pins.cpp -- TEST INFRASTRUCTURE (never linked into the product): the reference's OWN keypoint draw and descriptor
// cycle-matching statements (ReprojectionFactor's constructor, core/gtsam/reprojection_factor.cpp:42-89 -- the same statements
// stand in match_geometry_factor.cpp:62-97 and camera_tracker.cpp:798-834), #included from a file oracle/build_loop_ref.py extracts
// verbatim at build time (scratch directory, removed after the build), run with libtorch on the CPU.  This file supplies the names the block reads: kf_ / fr_ with an
// id and a descriptor map, valid_locations_1d, width, height, num_points, num_keypoints_, cyc_consis_thresh.
//   stdin:  C H W  num_keypoints  kf_id fr_id  thresh  N   then N valid locations, C*H*W values of desc0, C*H*W of desc1
//   stdout: "I n" keypoint_indexes_, "R n" raw_matched_locations_1d_1, "C n" cyc_matched_locations_1d_0, "M n" matched_keypoint_indexes_
#include <algorithm>
#include <cstdio>
#include <iostream>
#include <numeric>
#include <random>
#include <vector>

#include <torch/torch.h>

struct FrameStub
{
  long id;
  at::Tensor feat_desc;
};

static void dump(const char *tag, const at::Tensor t)
{
  const at::Tensor c = t.to(torch::kCPU).to(torch::kLong).reshape({-1}).contiguous();
  std::printf("%s %ld\n", tag, (long)c.numel());
  for (long i = 0; i < c.numel(); ++i)
    std::printf("%ld\n", c.data_ptr<long>()[i]);
}

int main()
{
  torch::NoGradGuard no_grad;
  using namespace torch::indexing;
  long channel_in, height, width, num_keypoints, kf_id, fr_id, N;
  float cyc_consis_thresh;
  if (!(std::cin >> channel_in >> height >> width >> num_keypoints >> kf_id >> fr_id >> cyc_consis_thresh >> N))
    return 2;
  std::vector<long> loc(N);
  for (long &v : loc)
    std::cin >> v;
  std::vector<float> d0(channel_in * height * width), d1(d0.size());
  for (float &v : d0)
    std::cin >> v;
  for (float &v : d1)
    std::cin >> v;
  FrameStub kf{kf_id, torch::from_blob(d0.data(), {1, channel_in, height, width}, torch::kFloat32).clone()};
  FrameStub fr{fr_id, torch::from_blob(d1.data(), {1, channel_in, height, width}, torch::kFloat32).clone()};
  FrameStub *kf_ = &kf, *fr_ = &fr;
  const at::Tensor valid_locations_1d = torch::from_blob(loc.data(), {N}, torch::kLong).clone();
  const long num_points = N;
  long num_keypoints_ = (num_points >= num_keypoints) ? num_keypoints : num_points; // :40
  at::Tensor keypoint_indexes_, matched_keypoint_indexes_;
#include "ref_match_block.h"
  dump("I", keypoint_indexes_);
  dump("R", raw_matched_locations_1d_1);
  dump("C", cyc_matched_locations_1d_0);
  dump("M", matched_keypoint_indexes_);
  return 0;
}

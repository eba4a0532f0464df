// producer_pins.cpp -- TEST INFRASTRUCTURE (never linked into the product): runs the reference's OWN input producers on the CPU.
//
// #included from files that oracle/build_loop_ref.py extracts verbatim from /root/reference at build time (scratch directory, removed after the build; nothing is
// copied into the repository):
//   * GenerateMaskPyramid            core/mapping/mapping_utils.cpp:321-342
//   * ComputeSpatialGrad             core/mapping/mapping_utils.h:236-256
//   * GenerateValidLocations         core/mapping/mapping_utils.h:258-296
//   * Mapper::GenerateGaussianPyramidWithGrad   core/mapping/mapper.cpp:1383-1426, and the Gaussian kernel / convolution options of
//     the Mapper constructor (:30-37; the `.to(torch::kCUDA, cuda_id_)` of that statement is dropped by the recipe: libtorch CPU here)
// This file supplies the members the extracted method reads (output_mask_pyramid_ptr_, gauss_kernel_, gauss_conv_options_) and a
// PinholeCamera<float> with the accessors GenerateValidLocations calls.
//   stdin:  F H W L  fx fy u0 v0   then H*W mask values, then F*H*W feature values
//   stdout: "M <n>" + level masks (all levels concatenated), "P <n>" + pyramid [F, SP], "G <n>" + gradient pyramid [2, F, SP],
//           "L <n>" + valid locations, "H <n>" + their homogeneous rays [N, 3]
#include <cstdio>
#include <iostream>
#include <memory>
#include <vector>

#include <torch/torch.h>

namespace df
{
template <typename T>
struct PinholeCamera
{
  T fx_, fy_, u0_, v0_, w_, h_;
  T fx() const { return fx_; }
  T fy() const { return fy_; }
  T u0() const { return u0_; }
  T v0() const { return v0_; }
  T width() const { return w_; }
  T height() const { return h_; }
};

#include "ref_producer_utils.h"

struct MapperStub
{
  std::shared_ptr<std::vector<at::Tensor>> output_mask_pyramid_ptr_;
  at::Tensor gauss_kernel_;
  torch::nn::functional::Conv2dFuncOptions gauss_conv_options_;
  torch::nn::functional::InterpolateFuncOptions mask_interp_options_;
  int cuda_id_ = 0;
  MapperStub()
  {
    namespace F = torch::nn::functional;
#include "ref_producer_kernel.h"
  }
#include "ref_producer_member.h"
};
} // namespace df

static void dump(const char *tag, const at::Tensor t)
{
  const at::Tensor c = t.to(torch::kFloat64).reshape({-1}).contiguous();
  std::printf("%s %ld\n", tag, (long)c.numel());
  const double *p = c.data_ptr<double>();
  for (long i = 0; i < c.numel(); ++i)
    std::printf("%.9g\n", p[i]);
}

int main()
{
  torch::NoGradGuard no_grad;
  long F, H, W, L;
  float fx, fy, u0, v0;
  if (!(std::cin >> F >> H >> W >> L >> fx >> fy >> u0 >> v0))
    return 2;
  std::vector<float> mask(H * W), feat(F * H * W);
  for (float &v : mask)
    std::cin >> v;
  for (float &v : feat)
    std::cin >> v;
  const at::Tensor m = torch::from_blob(mask.data(), {1, 1, H, W}, torch::kFloat32).clone();
  const at::Tensor f = torch::from_blob(feat.data(), {1, F, H, W}, torch::kFloat32).clone();
  df::MapperStub mapper;
  mapper.output_mask_pyramid_ptr_ = std::make_shared<std::vector<at::Tensor>>();
  df::GenerateMaskPyramid(m, (int)L, mapper.output_mask_pyramid_ptr_);
  std::vector<at::Tensor> flat;
  for (const at::Tensor &lvl : *mapper.output_mask_pyramid_ptr_)
    flat.push_back(lvl.reshape({-1}));
  dump("M", torch::cat(flat, 0));
  at::Tensor pyr, grad;
  mapper.GenerateGaussianPyramidWithGrad(f, (int)L, pyr, grad);
  dump("P", pyr);
  dump("G", grad);
  df::PinholeCamera<float> cam{fx, fy, u0, v0, (float)W, (float)H};
  at::Tensor norm2d, loc1d, homo;
  df::GenerateValidLocations(m, cam, norm2d, loc1d, homo);
  dump("L", loc1d);
  dump("H", homo);
  return 0;
}

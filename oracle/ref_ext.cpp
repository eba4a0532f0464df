// pybind11 front for the reference's own df::*_calculate functions (TEST INFRASTRUCTURE).
// This file is ours; the functions it calls are compiled, unmodified, from
// /root/reference/system/sources/cuda/*.cpp by oracle/build_ref.py.  It exists so that
// tests/golden/ can be generated from -- and bench.py's reference-GPU leg can time -- the
// reference itself on the same B200.
#include <torch/extension.h>
#include "photometric_factor_kernels.h"
#include "geometric_factor_kernels.h"
#include "reprojection_factor_kernels.h"
#include "match_geometry_factor_kernels.h"

#ifndef DF_CODE_SIZE
#error "DF_CODE_SIZE / DF_FEAT_SIZE must be defined"
#endif

namespace
{
constexpr int CS = DF_CODE_SIZE;
constexpr int FS = DF_FEAT_SIZE;

df::PinholeCamera<float> make_cam(const std::vector<double> &c)
{
  return df::PinholeCamera<float>((float)c[0], (float)c[1], (float)c[2], (float)c[3], (float)c[4], (float)c[5]);
}
df::CameraPyramid<float> make_pyr(const std::vector<double> &c, int levels)
{
  return df::CameraPyramid<float>(make_cam(c), (std::size_t)levels);
}

std::vector<std::vector<double>> camera_pyramid(const std::vector<double> &c, int levels)
{
  auto pyr = make_pyr(c, levels);
  std::vector<std::vector<double>> out;
  for (int i = 0; i < levels; ++i)
    out.push_back({pyr[i].fx(), pyr[i].fy(), pyr[i].u0(), pyr[i].v0(), pyr[i].width(), pyr[i].height()});
  return out;
}

std::tuple<at::Tensor, at::Tensor, double> photometric_jac_error(
    at::Tensor R10, at::Tensor t10, at::Tensor R0, at::Tensor t0, at::Tensor R1, at::Tensor t1,
    at::Tensor bias0, at::Tensor jac0, at::Tensor code0, at::Tensor mask1, at::Tensor loc1d, at::Tensor homo,
    at::Tensor feat0, at::Tensor feat1, at::Tensor grad1, at::Tensor level_offsets, double scale0,
    std::vector<double> cam, int levels, double eps, at::Tensor weights)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::photometric_jac_error_calculate<CS, FS>(AtA, Atb, err, R10, t10, R0, t0, R1, t1, bias0, jac0, code0, mask1,
                                              loc1d, homo, feat0, feat1, grad1, level_offsets, (float)scale0,
                                              make_pyr(cam, levels), (float)eps, weights);
  return {AtA, Atb, (double)err};
}

double photometric_error(at::Tensor R, at::Tensor t, at::Tensor bias0, at::Tensor jac0, at::Tensor code0,
                         at::Tensor mask1, at::Tensor loc1d, at::Tensor homo, at::Tensor feat0, at::Tensor feat1,
                         at::Tensor level_offsets, double scale0, std::vector<double> cam, int levels, double eps,
                         at::Tensor weights)
{
  return df::photometric_error_calculate<FS>(R, t, bias0, jac0, code0, mask1, loc1d, homo, feat0, feat1,
                                             level_offsets, (float)scale0, make_pyr(cam, levels), (float)eps, weights);
}

std::tuple<at::Tensor, at::Tensor, double> tracker_photo_jac_error(
    at::Tensor R, at::Tensor t, at::Tensor mask1, at::Tensor dpts0, at::Tensor homo, at::Tensor sfeat0,
    at::Tensor feat1, at::Tensor grad1, at::Tensor level_offsets, std::vector<double> cam, int levels, double eps,
    at::Tensor weights)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::tracker_photo_jac_error_calculate<FS>(AtA, Atb, err, R, t, mask1, dpts0, homo, sfeat0, feat1, grad1,
                                            level_offsets, make_pyr(cam, levels), (float)eps, weights);
  return {AtA, Atb, (double)err};
}

std::tuple<at::Tensor, at::Tensor, double> tracker_photo_jac_error_with_scale(
    at::Tensor R, at::Tensor t, at::Tensor mask1, at::Tensor dpts0, at::Tensor homo, at::Tensor sfeat0,
    at::Tensor feat1, at::Tensor grad1, at::Tensor level_offsets, std::vector<double> cam, int levels, double scale0,
    double eps, at::Tensor weights)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::tracker_photo_jac_error_calculate_with_scale<FS>(AtA, Atb, err, R, t, mask1, dpts0, homo, sfeat0, feat1, grad1,
                                                       level_offsets, make_pyr(cam, levels), (float)scale0, (float)eps,
                                                       weights);
  return {AtA, Atb, (double)err};
}

double tracker_photo_error(at::Tensor R, at::Tensor t, at::Tensor mask1, at::Tensor dpts0, at::Tensor homo,
                           at::Tensor sfeat0, at::Tensor feat1, at::Tensor level_offsets, std::vector<double> cam,
                           int levels, double eps, at::Tensor weights)
{
  return df::tracker_photo_error_calculate<FS>(R, t, mask1, dpts0, homo, sfeat0, feat1, level_offsets,
                                               make_pyr(cam, levels), (float)eps, weights);
}

std::tuple<at::Tensor, at::Tensor, double> geometric_jac_error(
    at::Tensor R10, at::Tensor t10, at::Tensor R0, at::Tensor t0, at::Tensor R1, at::Tensor t1, at::Tensor bias0,
    at::Tensor jac0, at::Tensor code0, at::Tensor dpt1, at::Tensor dgrad1, at::Tensor basis1, at::Tensor mask1,
    at::Tensor loc1d, at::Tensor homo, double scale0, double scale1, std::vector<double> cam, double eps,
    double loss_param, double weight)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::geometric_jac_error_calculate<CS>(AtA, Atb, err, R10, t10, R0, t0, R1, t1, bias0, jac0, code0, dpt1, dgrad1,
                                        basis1, mask1, loc1d, homo, (float)scale0, (float)scale1, make_cam(cam),
                                        (float)eps, (float)loss_param, (float)weight);
  return {AtA, Atb, (double)err};
}

double geometric_error(at::Tensor R, at::Tensor t, at::Tensor bias0, at::Tensor jac0, at::Tensor code0,
                       at::Tensor dpt1, at::Tensor mask1, at::Tensor loc1d, at::Tensor homo, double scale0,
                       std::vector<double> cam, double eps, double loss_param, double weight)
{
  return df::geometric_error_calculate<CS>(R, t, bias0, jac0, code0, dpt1, mask1, loc1d, homo, (float)scale0,
                                           make_cam(cam), (float)eps, (float)loss_param, (float)weight);
}

std::tuple<at::Tensor, at::Tensor, double> reprojection_jac_error(
    at::Tensor R10, at::Tensor t10, at::Tensor R0, at::Tensor t0, at::Tensor R1, at::Tensor t1, at::Tensor bias0,
    at::Tensor jac0, at::Tensor code0, at::Tensor loc1d, at::Tensor homo, at::Tensor match2d, double scale0,
    std::vector<double> cam, double eps, double loss_param, double weight)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::reprojection_jac_error_calculate<CS>(AtA, Atb, err, R10, t10, R0, t0, R1, t1, bias0, jac0, code0, loc1d, homo,
                                           match2d, (float)scale0, make_cam(cam), (float)eps, (float)loss_param,
                                           (float)weight);
  return {AtA, Atb, (double)err};
}

double reprojection_error(at::Tensor R, at::Tensor t, at::Tensor bias0, at::Tensor jac0, at::Tensor code0,
                          at::Tensor loc1d, at::Tensor homo, at::Tensor match2d, double scale0,
                          std::vector<double> cam, double eps, double loss_param, double weight)
{
  return df::reprojection_error_calculate<CS>(R, t, bias0, jac0, code0, loc1d, homo, match2d, (float)scale0,
                                              make_cam(cam), (float)eps, (float)loss_param, (float)weight);
}

std::tuple<at::Tensor, at::Tensor, double> tracker_reproj_jac_error(at::Tensor R, at::Tensor t, at::Tensor dpts0,
                                                                    at::Tensor homo, at::Tensor match2d,
                                                                    std::vector<double> cam, double eps,
                                                                    double loss_param, double weight)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::tracker_reproj_jac_error_calculate(AtA, Atb, err, R, t, dpts0, homo, match2d, make_cam(cam), (float)eps,
                                         (float)loss_param, (float)weight);
  return {AtA, Atb, (double)err};
}

double tracker_reproj_error(at::Tensor R, at::Tensor t, at::Tensor dpts0, at::Tensor homo, at::Tensor match2d,
                            std::vector<double> cam, double eps, double loss_param, double weight)
{
  return df::tracker_reproj_error_calculate(R, t, dpts0, homo, match2d, make_cam(cam), (float)eps, (float)loss_param,
                                            (float)weight);
}
std::tuple<at::Tensor, at::Tensor, double> tracker_match_geom_jac_error(at::Tensor R, at::Tensor t, at::Tensor dpts0, at::Tensor dpts1,
                                                                        at::Tensor homo0, at::Tensor homo1, double loss_param,
                                                                        double weight)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::tracker_match_geom_jac_error_calculate(AtA, Atb, err, R, t, dpts0, dpts1, homo0, homo1, (float)loss_param, (float)weight);
  return {AtA, Atb, (double)err};
}

std::tuple<at::Tensor, at::Tensor, double> tracker_match_geom_jac_error_with_scale(at::Tensor R, at::Tensor t, at::Tensor dpts0,
                                                                                   at::Tensor dpts1, at::Tensor homo0, at::Tensor homo1,
                                                                                   double scale0, double loss_param, double weight)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::tracker_match_geom_jac_error_calculate_with_scale(AtA, Atb, err, R, t, dpts0, dpts1, homo0, homo1, (float)scale0,
                                                        (float)loss_param, (float)weight);
  return {AtA, Atb, (double)err};
}

double tracker_match_geom_error(at::Tensor R, at::Tensor t, at::Tensor dpts0, at::Tensor dpts1, at::Tensor homo0, at::Tensor homo1,
                                double loss_param, double weight)
{
  return df::tracker_match_geom_error_calculate(R, t, dpts0, dpts1, homo0, homo1, (float)loss_param, (float)weight);
}
std::tuple<at::Tensor, at::Tensor, double> match_geometry_jac_error(
    at::Tensor R10, at::Tensor t10, at::Tensor R0, at::Tensor t0, at::Tensor R1, at::Tensor t1, at::Tensor bias0, at::Tensor bias1,
    at::Tensor jac0, at::Tensor jac1, at::Tensor code0, at::Tensor code1, at::Tensor homo0, at::Tensor homo1, at::Tensor loc0,
    at::Tensor loc1, double scale0, double scale1, double loss_param, double weight, std::string loss_type)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::match_geometry_jac_error_calculate<CS>(AtA, Atb, err, R10, t10, R0, t0, R1, t1, bias0, bias1, jac0, jac1, code0, code1, homo0,
                                             homo1, loc0, loc1, (float)scale0, (float)scale1, (float)loss_param, (float)weight,
                                             loss_type);
  return {AtA, Atb, (double)err};
}

double match_geometry_error(at::Tensor R, at::Tensor t, at::Tensor bias0, at::Tensor bias1, at::Tensor jac0, at::Tensor jac1,
                            at::Tensor code0, at::Tensor code1, at::Tensor homo0, at::Tensor homo1, at::Tensor loc0, at::Tensor loc1,
                            double scale0, double scale1, double loss_param, double weight, std::string loss_type)
{
  return df::match_geometry_error_calculate<CS>(R, t, bias0, bias1, jac0, jac1, code0, code1, homo0, homo1, loc0, loc1, (float)scale0,
                                                (float)scale1, (float)loss_param, (float)weight, loss_type);
}

std::tuple<at::Tensor, at::Tensor, double> loop_mg_jac_error(at::Tensor R10, at::Tensor t10, at::Tensor R0, at::Tensor t0, at::Tensor R1,
                                                             at::Tensor t1, at::Tensor dpts0, at::Tensor dpts1, at::Tensor homo0,
                                                             at::Tensor homo1, double scale0, double scale1, double loss_param,
                                                             double weight)
{
  at::Tensor AtA, Atb;
  float err = 0;
  df::loop_mg_jac_error_calculate(AtA, Atb, err, R10, t10, R0, t0, R1, t1, dpts0, dpts1, homo0, homo1, (float)scale0, (float)scale1,
                                  (float)loss_param, (float)weight);
  return {AtA, Atb, (double)err};
}

double loop_mg_error(at::Tensor R, at::Tensor t, at::Tensor dpts0, at::Tensor dpts1, at::Tensor homo0, at::Tensor homo1, double scale0,
                     double scale1, double loss_param, double weight)
{
  return df::loop_mg_error_calculate(R, t, dpts0, dpts1, homo0, homo1, (float)scale0, (float)scale1, (float)loss_param, (float)weight);
}
} // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
  m.attr("CODE_SIZE") = CS;
  m.attr("FEAT_SIZE") = FS;
  m.def("camera_pyramid", &camera_pyramid);
  m.def("photometric_jac_error", &photometric_jac_error);
  m.def("photometric_error", &photometric_error);
  m.def("tracker_photo_jac_error", &tracker_photo_jac_error);
  m.def("tracker_photo_jac_error_with_scale", &tracker_photo_jac_error_with_scale);
  m.def("tracker_photo_error", &tracker_photo_error);
  m.def("geometric_jac_error", &geometric_jac_error);
  m.def("geometric_error", &geometric_error);
  m.def("reprojection_jac_error", &reprojection_jac_error);
  m.def("reprojection_error", &reprojection_error);
  m.def("tracker_reproj_jac_error", &tracker_reproj_jac_error);
  m.def("tracker_reproj_error", &tracker_reproj_error);
  m.def("tracker_match_geom_jac_error", &tracker_match_geom_jac_error);
  m.def("tracker_match_geom_jac_error_with_scale", &tracker_match_geom_jac_error_with_scale);
  m.def("tracker_match_geom_error", &tracker_match_geom_error);
  m.def("match_geometry_jac_error", &match_geometry_jac_error);
  m.def("match_geometry_error", &match_geometry_error);
  m.def("loop_mg_jac_error", &loop_mg_jac_error);
  m.def("loop_mg_error", &loop_mg_error);
}

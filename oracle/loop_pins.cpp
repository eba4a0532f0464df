// loop_pins.cpp -- TEST INFRASTRUCTURE (never linked into the product): runs the reference's OWN tracker LM loops over a synthetic
// cost and prints every cost evaluation they make.
//
// The loops -- CameraTracker::TrackNewFrame (core/system/camera_tracker.cpp:1156-1279, 6-DoF) and CameraTracker::TrackFrame
// (:1479-1630, 7-DoF) -- their declaration blocks, UpdateVariables (:467-512) and LMConvergence (:527-573) -- and UpdateDepth
// (core/mapping/mapping_utils.h:216-222, second mode below) are #included from files
// that oracle/build_loop_ref.py extracts verbatim from /root/reference at build time (into a scratch directory removed after the build; nothing is copied into the
// repository).  This file only supplies what those pieces refer to: the members config_ / tracker_name_ / kf_ / photo_weights_tensor_,
// a VLOG sink, and ComputeJacobianAndError / ComputeError over a pinhole reprojection cost on a point cloud
//     x_i = R (s p_i) + t,   r_i = [pi(x_i) - uv_i ; wz (x_i.z - d_i)],   error = mean |r_i|^2,   AtA = mean J^T J,  Atb = -mean J^T r,
//     J_i = d r / d x [ I | -[x_i]x | R p_i ]   (left-multiplicative increment, the convention of UpdateVariables; the depth row,
//     weight wz = 0 in the 6-DoF cases, fixes the scale gauge of the 7-DoF ones like the reference's match-geometry factor does),
// evaluated in double and rounded to float.  oracle/make_golden_loop.py feeds cases on stdin and stores the evaluation log in
// tests/golden/loop_pins.npz; tests/test_loop_pins.py holds oracle.tracker_lm / tracker_lm7 (the restatement the GPU tracker is
// tested against) to that log, evaluation by evaluation.
//
//   stdin:  dof(6|7) n  init_damp min_damp max_damp damp_inc damp_dec jac_thresh min_grad min_param_inc max_iters  fx fy cx cy
//           wz  R(9) t(3) s  then n lines: px py pz u v d
//   stdout: one line per evaluation  "J|E  R(9) t(3) s  error"  then  "F  R(9) t(3) s  curr_error  curr_iter"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <tuple>
#include <vector>

#include <Eigen/Dense>
#include <torch/torch.h>

struct NullLog
{
  template <class T>
  NullLog &operator<<(const T &) { return *this; }
};
#define VLOG(n) NullLog()
namespace cv
{
struct Mat
{
};
} // namespace cv

namespace df
{
#include "ref_loop_utils.h"

struct Config
{
  float init_damp, min_damp, max_damp, damp_inc_factor, damp_dec_factor, jac_update_err_inc_threshold, min_grad_thresh, min_param_inc_thresh;
  long max_num_iters;
};
struct FrameStub
{
  long id = 0;
};

struct Cost
{
  double fx, fy, cx, cy, wz = 0.0;
  std::vector<std::array<double, 6>> pts; // px py pz u v d
  // error only when A == nullptr
  double eval(const double *R, const double *t, double s, double *A, double *b, int dof) const
  {
    if (A)
    {
      std::memset(A, 0, sizeof(double) * dof * dof);
      std::memset(b, 0, sizeof(double) * dof);
    }
    double err = 0.0;
    for (const auto &p : pts)
    {
      const double q[3] = {s * p[0], s * p[1], s * p[2]};
      double Rp[3], x[3];
      for (int k = 0; k < 3; ++k)
      {
        Rp[k] = R[k * 3 + 0] * p[0] + R[k * 3 + 1] * p[1] + R[k * 3 + 2] * p[2];
        x[k] = R[k * 3 + 0] * q[0] + R[k * 3 + 1] * q[1] + R[k * 3 + 2] * q[2] + t[k];
      }
      const double iz = 1.0 / x[2];
      const double r[3] = {fx * x[0] * iz + cx - p[3], fy * x[1] * iz + cy - p[4], wz * (x[2] - p[5])};
      err += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      if (A)
      {
        const double P[3][3] = {{fx * iz, 0.0, -fx * x[0] * iz * iz}, {0.0, fy * iz, -fy * x[1] * iz * iz}, {0.0, 0.0, wz}};
        // d x / d [v | w | s] = [ I | -[x]x | R p ]
        const double D[3][7] = {{1, 0, 0, 0, x[2], -x[1], Rp[0]}, {0, 1, 0, -x[2], 0, x[0], Rp[1]}, {0, 0, 1, x[1], -x[0], 0, Rp[2]}};
        double J[3][7];
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 7; ++c)
            J[a][c] = P[a][0] * D[0][c] + P[a][1] * D[1][c] + P[a][2] * D[2][c];
        for (int i = 0; i < dof; ++i)
        {
          for (int j = 0; j < dof; ++j)
            A[i * dof + j] += J[0][i] * J[0][j] + J[1][i] * J[1][j] + J[2][i] * J[2][j];
          b[i] -= J[0][i] * r[0] + J[1][i] * r[1] + J[2][i] * r[2];
        }
      }
    }
    const double inv = 1.0 / (double)pts.size();
    if (A)
    {
      for (int i = 0; i < dof * dof; ++i)
        A[i] *= inv;
      for (int i = 0; i < dof; ++i)
        b[i] *= inv;
    }
    return err * inv;
  }
};

struct LoopHarness
{
  Config config_;
  std::string tracker_name_ = "loop_pins";
  FrameStub kf_storage_, *kf_ = &kf_storage_;
  at::Tensor photo_weights_tensor_ = torch::ones({4});
  Cost cost;

  static void pose_of(const at::Tensor R, const at::Tensor t, double *Rd, double *td)
  {
    const at::Tensor Rc = R.to(torch::kCPU).to(torch::kFloat64).contiguous(), tc = t.to(torch::kCPU).to(torch::kFloat64).reshape({-1}).contiguous();
    std::memcpy(Rd, Rc.data_ptr<double>(), sizeof(double) * 9);
    std::memcpy(td, tc.data_ptr<double>(), sizeof(double) * 3);
  }
  static void log(char kind, const double *R, const double *t, double s, float error)
  {
    std::printf("%c", kind);
    for (int i = 0; i < 9; ++i)
      std::printf(" %.9g", (double)(float)R[i]);
    for (int i = 0; i < 3; ++i)
      std::printf(" %.9g", (double)(float)t[i]);
    std::printf(" %.9g %.9g\n", (double)(float)s, (double)error);
  }
  void jac(const at::Tensor R, const at::Tensor t, double s, int dof, bool update_error, at::Tensor &AtA, at::Tensor &Atb, float &error)
  {
    double Rd[9], td[3], A[49], b[7];
    pose_of(R, t, Rd, td);
    const float e = (float)cost.eval(Rd, td, s, A, b, dof);
    AtA = torch::from_blob(A, {dof, dof}, torch::kFloat64).to(torch::kFloat32).clone();
    Atb = torch::from_blob(b, {dof, 1}, torch::kFloat64).to(torch::kFloat32).clone(); // [D, 1] like the reference kernels
    if (update_error)
      error = e;
    log('J', Rd, td, s, e);
  }
  void err(const at::Tensor R, const at::Tensor t, double s, float &error)
  {
    double Rd[9], td[3];
    pose_of(R, t, Rd, td);
    error = (float)cost.eval(Rd, td, s, nullptr, nullptr, 6);
    log('E', Rd, td, s, error);
  }
  // the loops call these with the data tensors of the real factors first; only the state and the outputs matter here
  template <class... Ts>
  void ComputeJacobianAndError(Ts &&...args)
  {
    auto a = std::forward_as_tuple(args...);
    constexpr int n = sizeof...(Ts);
    if constexpr (n == 15) // ..., R, t, use_photo, use_reproj, update_error, AtA, Atb, error
      jac(std::get<7>(a), std::get<8>(a), 1.0, 6, std::get<11>(a), std::get<12>(a), std::get<13>(a), std::get<14>(a));
    else // ..., R, t, scale, use_photo, use_match_geom, update_error, AtA, Atb, error
      jac(std::get<8>(a), std::get<9>(a), (double)std::get<10>(a), 7, std::get<13>(a), std::get<14>(a), std::get<15>(a), std::get<16>(a));
  }
  template <class... Ts>
  void ComputeError(Ts &&...args)
  {
    auto a = std::forward_as_tuple(args...);
    constexpr int n = sizeof...(Ts);
    if constexpr (n == 12) // ..., R, t, use_photo, use_reproj, error
      err(std::get<7>(a), std::get<8>(a), 1.0, std::get<11>(a));
    else // ..., R, t, scale, use_photo, use_match_geom, error
      err(std::get<8>(a), std::get<9>(a), (double)std::get<10>(a), std::get<n - 1>(a));
  }

#include "ref_loop_members.h"

  bool TrackNewFrameLoop(at::Tensor guess_rotation_10, at::Tensor guess_translation_10)
  {
    FrameStub frame_to_track;
    at::Tensor cat_photo_features_0, photo_dpts_0, photo_locations_homo_0, inlier_keypoint_dpts_0, inlier_keypoint_locations_homo_0,
        matched_locations_2d_1;
    const bool use_photo = true, use_reproj = true;
#include "ref_loop_decl6.h"
#include "ref_loop_body6.h"
    double Rd[9], td[3];
    pose_of(guess_rotation_10, guess_translation_10, Rd, td);
    std::printf("F");
    for (int i = 0; i < 9; ++i)
      std::printf(" %.9g", (double)(float)Rd[i]);
    for (int i = 0; i < 3; ++i)
      std::printf(" %.9g", (double)(float)td[i]);
    std::printf(" 1 %.9g %ld\n", (double)curr_error, curr_iter);
    return true;
  }

  bool TrackFrameLoop(at::Tensor guess_rotation, at::Tensor guess_translation, float guess_scale)
  {
    FrameStub frame_to_track;
    at::Tensor cat_photo_features_0, unscaled_photo_dpts_0, photo_locations_homo_0, unscaled_inlier_keypoint_dpts_0,
        inlier_keypoint_locations_homo_0, matched_dpts_1, matched_locations_homo_1;
    const bool use_photo = true, use_match_geom = true;
#include "ref_loop_decl7.h"
#include "ref_loop_rest7.h"
#include "ref_loop_body7.h"
    double Rd[9], td[3];
    pose_of(guess_rotation, guess_translation, Rd, td);
    std::printf("F");
    for (int i = 0; i < 9; ++i)
      std::printf(" %.9g", (double)(float)Rd[i]);
    for (int i = 0; i < 3; ++i)
      std::printf(" %.9g", (double)(float)td[i]);
    std::printf(" %.9g %.9g %ld\n", (double)guess_scale, (double)curr_error, curr_iter);
    return true;
  }
};
} // namespace df

// "0 H W C scale" then H*W bias values, H*W*C basis values (row = pixel), C code values -> the H*W values UpdateDepth writes
static int update_depth_mode()
{
  long H, W, C;
  float scale;
  std::cin >> H >> W >> C >> scale;
  std::vector<float> bias(H * W), jac(H * W * C), code(C);
  for (float &v : bias)
    std::cin >> v;
  for (float &v : jac)
    std::cin >> v;
  for (float &v : code)
    std::cin >> v;
  const at::Tensor b = torch::from_blob(bias.data(), {1, 1, H, W}, torch::kFloat32), J = torch::from_blob(jac.data(), {H * W, C}, torch::kFloat32),
                   c = torch::from_blob(code.data(), {C}, torch::kFloat32);
  at::Tensor dpt_map;
  df::UpdateDepth<float>(b, J, c, scale, dpt_map);
  const at::Tensor out = dpt_map.reshape({-1}).contiguous();
  for (long i = 0; i < H * W; ++i)
    std::printf("%.9g\n", (double)out.data_ptr<float>()[i]);
  return 0;
}

int main()
{
  torch::NoGradGuard no_grad;
  df::LoopHarness h;
  int dof = 0, n = 0;
  auto &c = h.config_;
  if (!(std::cin >> dof))
    return 2;
  if (dof == 0)
    return update_depth_mode();
  if (!(std::cin >> n >> c.init_damp >> c.min_damp >> c.max_damp >> c.damp_inc_factor >> c.damp_dec_factor >> c.jac_update_err_inc_threshold >>
        c.min_grad_thresh >> c.min_param_inc_thresh >> c.max_num_iters >> h.cost.fx >> h.cost.fy >> h.cost.cx >> h.cost.cy >> h.cost.wz))
    return 2;
  float R[9], t[3], s;
  for (float &v : R)
    std::cin >> v;
  for (float &v : t)
    std::cin >> v;
  std::cin >> s;
  h.cost.pts.resize(n);
  for (auto &p : h.cost.pts)
    for (double &v : p)
      std::cin >> v;
  at::Tensor Rt = torch::from_blob(R, {3, 3}, torch::kFloat32).clone(), tt = torch::from_blob(t, {3, 1}, torch::kFloat32).clone();
  if (dof == 6)
    h.TrackNewFrameLoop(Rt, tt);
  else
    h.TrackFrameLoop(Rt, tt, s);
  return 0;
}

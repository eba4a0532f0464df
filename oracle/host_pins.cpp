// host_pins.cpp -- golden vectors for the reference's host-side Eigen helpers, produced by the reference's OWN code.
//
// Test infrastructure (never linked into the product).  oracle/build_host_ref.py extracts, at build time and only where
// /root/reference exists, the template functions IsPsd / NearestPsd (core/mapping/mapping_utils.h:88-128) and so3_hat / se3_exp
// (:304-346) verbatim into oracle/_ref/ref_host_extract.h -- the header itself cannot be included (it pulls in GTSAM, OpenCV,
// torch, glog) -- and compiles this file against the Eigen 3.3.9 vendored by the reference (system/thirdparty/eigen).  The
// tracker's linear solve is the single expression of core/system/camera_tracker.cpp:1182-1183,
//     eigen_damped_AtA = eigen_AtA + curr_damp * eigen_AtA_diag;  eigen_solution = eigen_damped_AtA.colPivHouseholderQr().solve(eigen_Atb);
// in float, instantiated here for 6x6 (TrackNewFrame) and 7x7 (TrackFrame, :1523-1524).
//
// Protocol: reads a little-endian binary stream of cases from stdin, writes results to stdout.
//   case 'P' n  M[n*n] (double, row-major)            -> NearestPsd(M)                     n*n doubles
//   case 'E'    omega[3] v[3] (double)                -> se3_exp<double>: R[9] t[3]        12 doubles
//   case 'F'    omega[3] v[3] (float)                 -> se3_exp<float>:  R[9] t[3]        12 floats
//   case 'Q' n  AtA[n*n] Atb[n] damp (float), n = 6|7 -> colPivHouseholderQr solve         n floats
#include <Eigen/Dense>
#include <Eigen/SVD>
#include <Eigen/Eigenvalues>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

namespace df
{
#include "ref_host_extract.h"
}

template <typename T>
static bool rd(T *p, size_t n) { return fread(p, sizeof(T), n, stdin) == n; }
template <typename T>
static void wr(const T *p, size_t n) { fwrite(p, sizeof(T), n, stdout); }

template <int N>
static void qr_case(const std::vector<float> &A, const std::vector<float> &b, float damp)
{
  Eigen::Matrix<float, N, N> eigen_AtA, eigen_AtA_diag = Eigen::Matrix<float, N, N>::Zero(), eigen_damped_AtA;
  Eigen::Matrix<float, N, 1> eigen_Atb, eigen_solution;
  for (int r = 0; r < N; ++r)
  {
    for (int c = 0; c < N; ++c)
      eigen_AtA(r, c) = A[r * N + c];
    eigen_AtA_diag(r, r) = A[r * N + r]; // torch::diag(torch::diag(AtA)), camera_tracker.cpp:1171
    eigen_Atb(r) = b[r];
  }
  eigen_damped_AtA = eigen_AtA + damp * eigen_AtA_diag;
  eigen_solution = eigen_damped_AtA.colPivHouseholderQr().solve(eigen_Atb);
  wr(eigen_solution.data(), N);
}

int main()
{
  char tag;
  while (rd(&tag, 1))
  {
    if (tag == 'P')
    {
      int32_t n;
      rd(&n, 1);
      std::vector<double> m((size_t)n * n);
      rd(m.data(), m.size());
      Eigen::MatrixXd M(n, n);
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
          M(r, c) = m[(size_t)r * n + c];
      const Eigen::MatrixXd P = df::NearestPsd(M);
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
          m[(size_t)r * n + c] = P(r, c);
      wr(m.data(), m.size());
    }
    else if (tag == 'E' || tag == 'F')
    {
      if (tag == 'E')
      {
        double in[6], out[12];
        rd(in, 6);
        Eigen::Matrix<double, 3, 1> w(in[0], in[1], in[2]), v(in[3], in[4], in[5]), t;
        Eigen::Matrix<double, 3, 3> R;
        df::se3_exp<double>(w, v, R, t);
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            out[r * 3 + c] = R(r, c);
        for (int r = 0; r < 3; ++r)
          out[9 + r] = t(r);
        wr(out, 12);
      }
      else
      {
        float in[6], out[12];
        rd(in, 6);
        Eigen::Matrix<float, 3, 1> w(in[0], in[1], in[2]), v(in[3], in[4], in[5]), t;
        Eigen::Matrix<float, 3, 3> R;
        df::se3_exp<float>(w, v, R, t);
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            out[r * 3 + c] = R(r, c);
        for (int r = 0; r < 3; ++r)
          out[9 + r] = t(r);
        wr(out, 12);
      }
    }
    else if (tag == 'Q')
    {
      int32_t n;
      rd(&n, 1);
      std::vector<float> A((size_t)n * n), b(n);
      float damp;
      rd(A.data(), A.size());
      rd(b.data(), b.size());
      rd(&damp, 1);
      if (n == 6)
        qr_case<6>(A, b, damp);
      else if (n == 7)
        qr_case<7>(A, b, damp);
      else
        return 2;
    }
    else
      return 3;
  }
  return 0;
}

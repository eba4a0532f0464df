// retract_pins.cpp -- TEST INFRASTRUCTURE (never linked into the product): the reference's OWN SE(3) manifold for the mapping
// variables, gtsam::traits<Sophus::SE3<Scalar>> (core/gtsam/gtsam_traits.h:13-138: Retract, Local, its se3_exp), #included from a file
// that oracle/build_host_ref.py extracts verbatim at build time (scratch directory, removed after the build) and compiled against the Sophus and Eigen the reference
// vendors.  This file supplies the two GTSAM names the struct mentions (the primary template and gtsam::Vector).
//   stdin:  n, then n lines: R(9 row-major) t(3) delta(6: v, omega)
//   stdout: per line: float Retract -> R'(9) t'(3), double Retract -> R'(9) t'(3), Local(pose, Retract(pose, delta)) (6, float instantiation)
#include <cstdio>
#include <iostream>
#include <string>

#include <Eigen/Dense>
#include <sophus/se3.hpp>

namespace gtsam
{
template <typename T>
struct traits;
using Vector = Eigen::VectorXd;
} // namespace gtsam

namespace gtsam
{
#include "ref_se3_traits.h"
} // namespace gtsam

template <typename S>
static Sophus::SE3<S> make_pose(const double *R, const double *t)
{
  Eigen::Matrix<S, 3, 3> Rm;
  for (int i = 0; i < 9; ++i)
    Rm(i / 3, i % 3) = (S)R[i];
  // project on SO(3) the way a pose enters the reference's graph (Sophus stores a unit quaternion)
  Eigen::Quaternion<S> q(Rm);
  q.normalize();
  return Sophus::SE3<S>(Sophus::SO3<S>(q), Eigen::Matrix<S, 3, 1>((S)t[0], (S)t[1], (S)t[2]));
}

template <typename S>
static void print_pose(const Sophus::SE3<S> &p)
{
  const Eigen::Matrix<S, 3, 3> R = p.so3().matrix();
  for (int i = 0; i < 9; ++i)
    std::printf("%.17g ", (double)R(i / 3, i % 3));
  for (int i = 0; i < 3; ++i)
    std::printf("%.17g ", (double)p.translation()(i));
}

int main()
{
  int n;
  if (!(std::cin >> n))
    return 2;
  for (int k = 0; k < n; ++k)
  {
    double R[9], t[3];
    gtsam::Vector delta(6);
    for (double &v : R)
      std::cin >> v;
    for (double &v : t)
      std::cin >> v;
    for (int i = 0; i < 6; ++i)
      std::cin >> delta(i);
    const Sophus::SE3f pf = make_pose<float>(R, t);
    const Sophus::SE3d pd = make_pose<double>(R, t);
    const Sophus::SE3f qf = gtsam::traits<Sophus::SE3f>::Retract(pf, delta);
    const Sophus::SE3d qd = gtsam::traits<Sophus::SE3d>::Retract(pd, delta);
    print_pose(pf);
    print_pose(qf);
    print_pose(qd);
    const gtsam::Vector back = gtsam::traits<Sophus::SE3f>::Local(pf, qf);
    for (int i = 0; i < 6; ++i)
      std::printf("%.17g ", back(i));
    std::printf("\n");
  }
  return 0;
}

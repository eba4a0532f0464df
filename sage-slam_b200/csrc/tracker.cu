// tracker.cu -- CameraTracker::TrackNewFrame's damped Gauss-Newton loop on the 6-DoF relative pose
// (core/system/camera_tracker.cpp:1034-1310; loop :1156-1279, UpdateVariables :491-512,
// LMConvergence :552-573).  Same control flow and constants as the reference; the Jacobian / error
// evaluations are the fused kernels of photometric.cu / reprojection.cu and the keyframe features are
// pre-sampled once on the device (:1104-1123).  Host C++ above the C ABI.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "sage_internal.h"

using namespace sage;

namespace
{

// x = (A).colPivHouseholderQr().solve(b) for n <= 8 in float -- the reference's solve of the damped 6x6 / 7x7 system
// (core/system/camera_tracker.cpp:1182-1183, :1523-1524: Eigen::ColPivHouseholderQR<Matrix<float,n,n>>).  The algorithm of
// Eigen 3.3.9 (QR/ColPivHouseholderQR.h computeInPlace / _solve_impl, Householder/Householder.h makeHouseholderInPlace) is
// followed step by step -- column pivoting on the down-dated column norms with LAPACK's re-computation rule, the "nonzero
// pivots" cut (exact-zero sense, NOT the fuzzy rank), zeros for the columns beyond it -- so that nearly rank-deficient systems
// (textureless frames, damping at its floor) give the reference's step, not just well-conditioned ones.  Pinned by
// tests/golden/host_pins.npz (produced by the reference's expression compiled against its vendored Eigen).
bool solve_small(const float *A, const float *b, int n, float *x)
{
  float qr[8][8], hc[8], norm_upd[8], norm_dir[8], c[8];
  int perm[8];
  const float eps = 1.1920929e-07f, tiny = 1.17549435e-38f;
  for (int r = 0; r < n; ++r)
    for (int k = 0; k < n; ++k)
      qr[r][k] = A[r * n + k];
  float maxnorm = 0.f;
  for (int k = 0; k < n; ++k)
  {
    float s = 0.f;
    for (int r = 0; r < n; ++r)
      s += qr[r][k] * qr[r][k];
    norm_dir[k] = norm_upd[k] = std::sqrt(s);
    maxnorm = std::max(maxnorm, norm_upd[k]);
    perm[k] = k;
  }
  const float threshold_helper = (maxnorm * eps) * (maxnorm * eps) / (float)n;
  const float downdate_threshold = std::sqrt(eps);
  int nonzero = n;
  for (int k = 0; k < n; ++k)
  {
    int big = k;
    for (int j = k + 1; j < n; ++j)
      if (norm_upd[j] > norm_upd[big])
        big = j;
    if (nonzero == n && norm_upd[big] * norm_upd[big] < threshold_helper * (float)(n - k))
      nonzero = k;
    if (big != k)
    {
      for (int r = 0; r < n; ++r)
        std::swap(qr[r][k], qr[r][big]);
      std::swap(norm_upd[k], norm_upd[big]);
      std::swap(norm_dir[k], norm_dir[big]);
      std::swap(perm[k], perm[big]);
    }
    // Householder vector of column k (rows k..n-1): essential part stored below the diagonal, beta on it
    float tail2 = 0.f;
    for (int r = k + 1; r < n; ++r)
      tail2 += qr[r][k] * qr[r][k];
    const float c0 = qr[k][k];
    float tau, beta;
    if (tail2 <= tiny)
    {
      tau = 0.f;
      beta = c0;
      for (int r = k + 1; r < n; ++r)
        qr[r][k] = 0.f;
    }
    else
    {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0.f)
        beta = -beta;
      for (int r = k + 1; r < n; ++r)
        qr[r][k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    hc[k] = tau;
    qr[k][k] = beta;
    for (int j = k + 1; j < n; ++j) // apply H_k to the trailing columns
    {
      float tmp = qr[k][j];
      for (int r = k + 1; r < n; ++r)
        tmp += qr[r][k] * qr[r][j];
      qr[k][j] -= tau * tmp;
      for (int r = k + 1; r < n; ++r)
        qr[r][j] -= tau * qr[r][k] * tmp;
    }
    for (int j = k + 1; j < n; ++j) // norm down-date (LAPACK xGEQPF)
      if (norm_upd[j] != 0.f)
      {
        float t = std::fabs(qr[k][j]) / norm_upd[j];
        t = (1.f + t) * (1.f - t);
        t = t < 0.f ? 0.f : t;
        const float ratio = norm_upd[j] / norm_dir[j];
        if (t * ratio * ratio <= downdate_threshold)
        {
          float s = 0.f;
          for (int r = k + 1; r < n; ++r)
            s += qr[r][j] * qr[r][j];
          norm_dir[j] = norm_upd[j] = std::sqrt(s);
        }
        else
          norm_upd[j] *= std::sqrt(t);
      }
  }
  for (int i = 0; i < n; ++i)
    x[i] = 0.f;
  if (nonzero == 0)
    return true;
  for (int r = 0; r < n; ++r)
    c[r] = b[r];
  for (int k = 0; k < nonzero; ++k) // c = Q^T b = H_{r-1} ... H_0 b
  {
    float tmp = c[k];
    for (int r = k + 1; r < n; ++r)
      tmp += qr[r][k] * c[r];
    c[k] -= hc[k] * tmp;
    for (int r = k + 1; r < n; ++r)
      c[r] -= hc[k] * qr[r][k] * tmp;
  }
  for (int r = nonzero - 1; r >= 0; --r) // R y = c on the leading nonzero x nonzero triangle
  {
    float s = c[r];
    for (int k = r + 1; k < nonzero; ++k)
      s -= qr[r][k] * c[k];
    c[r] = s / qr[r][r];
  }
  for (int i = 0; i < nonzero; ++i)
    x[perm[i]] = c[i];
  for (int i = 0; i < n; ++i)
    if (!std::isfinite(x[i]))
      return false;
  return true;
}

void se3_exp_host(const float *w, const float *v, float *R, float *t)
{
  // se3_exp<float> (core/mapping/mapping_utils.h:316-346)
  float theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  float n[3] = {1.f, 0.f, 0.f};
  if (theta > 0.f)
    for (int i = 0; i < 3; ++i)
      n[i] = w[i] / theta;
  theta = std::max(theta, 1.0e-14f);
  const float s = std::sin(theta), c = std::cos(theta);
  const float K[9] = {0.f, -n[2], n[1], n[2], 0.f, -n[0], -n[1], n[0], 0.f};
  float K2[9];
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q)
      K2[r * 3 + q] = K[r * 3] * K[q] + K[r * 3 + 1] * K[3 + q] + K[r * 3 + 2] * K[6 + q];
  const float a = (1.0f - c) / theta, b = (theta - s) / theta;
  for (int r = 0; r < 3; ++r)
  {
    float tv = 0.f;
    for (int q = 0; q < 3; ++q)
    {
      const float id = r == q ? 1.f : 0.f;
      R[r * 3 + q] = id + s * K[r * 3 + q] + (1.0f - c) * K2[r * 3 + q];
      tv += (id + a * K[r * 3 + q] + b * K2[r * 3 + q]) * v[q];
    }
    t[r] = tv;
  }
}

// RotationToAngleAxis (core/mapping/mapping_utils.h:145-214), including its normalisation quirk
void rotation_to_angle_axis(const float *R, float eps, float *out)
{
  // m = R^T
  auto m = [&](int r, int c) { return R[c * 3 + r]; };
  const bool d2 = m(2, 2) < eps;
  const bool d0_d1 = m(0, 0) > m(1, 1);
  const bool d0_nd1 = m(0, 0) < -m(1, 1);
  const float t0 = 1.f + m(0, 0) - m(1, 1) - m(2, 2);
  const float q0[4] = {m(1, 2) - m(2, 1), t0, m(0, 1) + m(1, 0), m(2, 0) + m(0, 2)};
  const float t1 = 1.f - m(0, 0) + m(1, 1) - m(2, 2);
  const float q1[4] = {m(2, 0) - m(0, 2), m(0, 1) + m(1, 0), t1, m(1, 2) + m(2, 1)};
  const float t2 = 1.f - m(0, 0) - m(1, 1) + m(2, 2);
  const float q2[4] = {m(0, 1) - m(1, 0), m(2, 0) + m(0, 2), m(1, 2) + m(2, 1), t2};
  const float t3 = 1.f + m(0, 0) + m(1, 1) + m(2, 2);
  const float q3[4] = {t3, m(1, 2) - m(2, 1), m(2, 0) - m(0, 2), m(0, 1) - m(1, 0)};
  const float c0 = d2 && d0_d1, c1 = d2 && !d0_d1, c2 = !d2 && d0_nd1, c3 = !d2 && !d0_nd1;
  const float den = std::sqrt(t0 * c1 + t1 * c1 + t2 * c2 + t3 * c3); // reference: t0 * mask_c1 (:187)
  float q[4];
  for (int i = 0; i < 4; ++i)
    q[i] = 0.5f * (q0[i] * c0 + q1[i] * c1 + q2[i] * c2 + q3[i] * c3) / den;
  const float s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const float sn = std::sqrt(s2);
  const float two_theta = q[0] < 0.f ? std::atan2(-sn, -q[0]) : std::atan2(sn, q[0]);
  const float k = s2 > 0.f ? two_theta / sn : 2.0f;
  for (int i = 0; i < 3; ++i)
    out[i] = k * q[1 + i];
}

// The damped Gauss-Newton loop of CameraTracker::TrackNewFrame (core/system/camera_tracker.cpp:1156-1279, DOF = 6: relative pose)
// and CameraTracker::TrackFrame (:1479-1630, DOF = 7: + depth scale of the tracked frame), statement by statement: the Jacobian
// is recomputed only when the relative error change is large enough (:1159), the error of the linearisation is taken on the
// first iteration only (:1167), the damped system is solved with Eigen's colPivHouseholderQr in float (solve_small), the step
// is tested with LMConvergence (:527-573: max |Atb| and the SIGNED max of delta / (|x| + 1e-8)), applied with UpdateVariables
// (:467-512: T <- exp([v, w]) T, scale += delta[6]) and retried with more damping until it lowers the error or the damping is at
// its maximum.  jac(R, t, s, AtA, Atb, &err), err(R, t, s) -> error.  no_overlap_error >= 0 (TrackFrame without the match-geometry
// factor, :1500-1504): stop with return code 2 when the error says that nothing overlaps.
// The same code runs behind sage_ba_track_new_frame / sage_ba_track_frame (cost = the GPU factor kernels) and behind
// sage_ba_tracker_lm_callbacks (cost = the caller's), which is how tests/test_loop_pins.py holds it to the reference's own loops.
template <int DOF, class JacFn, class ErrFn>
static int lm_loop(const sage_ba_tracker_config *cfg, float *R, float *t, float *scale, float no_overlap_error, JacFn &&jac_fn, ErrFn &&err_fn,
                   sage_ba_tracker_report &rep)
{
  static_assert(DOF == 6 || DOF == 7, "relative pose (+ scale)");
  auto clampd = [&](float d) { return std::min(std::max(cfg->min_damp, d), cfg->max_damp); };
  float Rg[9], tg[3], Rc[9], tc[3], sg = scale ? *scale : 1.f, sc = sg;
  memcpy(Rg, R, sizeof(Rg));
  memcpy(tg, t, sizeof(tg));
  float AtA[DOF * DOF], Atb[DOF], damped[DOF * DOF], sol[DOF];
  float prev_error = 0.f, curr_error = 1.f, cand_error = 1.f; // :1056-1057
  float damp = cfg->init_damp;
  long iter = 0;
  bool update_jac = true;
  int rc = 0;
  while (true)
  {
    if (std::fabs(curr_error - prev_error) / prev_error > cfg->jac_update_err_inc_threshold)
    {
      float e = 0.f;
      jac_fn(Rg, tg, sg, AtA, Atb, &e);
      rep.jacobian_evals++;
      if (iter == 0)
        curr_error = e;
      update_jac = true;
    }
    else
      update_jac = false;
    if (no_overlap_error >= 0.f && curr_error >= no_overlap_error)
    {
      rc = 2;
      break;
    }
    iter += 1;
    auto solve = [&]() {
      for (int i = 0; i < DOF * DOF; ++i)
        damped[i] = AtA[i];
      for (int i = 0; i < DOF; ++i)
        damped[i * DOF + i] = AtA[i * DOF + i] + damp * AtA[i * DOF + i];
      if (!solve_small(damped, Atb, DOF, sol))
        for (int i = 0; i < DOF; ++i)
          sol[i] = 0.f;
    };
    solve();
    float rotvec[3];
    rotation_to_angle_axis(Rg, 1.0e-6f, rotvec);
    float max_grad = 0.f, max_inc = -INFINITY;
    for (int i = 0; i < DOF; ++i)
    {
      max_grad = std::max(max_grad, std::fabs(Atb[i]));
      const float x = i < 3 ? tg[i] : (i < 6 ? rotvec[i - 3] : sg);
      max_inc = std::max(max_inc, sol[i] / (std::fabs(x) + 1.0e-8f));
    }
    if (max_grad < cfg->min_grad_thresh || max_inc < cfg->min_param_inc_thresh)
      break;
    while (true)
    {
      float dR[9], dt[3];
      se3_exp_host(sol + 3, sol, dR, dt);
      for (int r = 0; r < 3; ++r)
      {
        for (int c = 0; c < 3; ++c)
          Rc[r * 3 + c] = dR[r * 3] * Rg[c] + dR[r * 3 + 1] * Rg[3 + c] + dR[r * 3 + 2] * Rg[6 + c];
        tc[r] = dR[r * 3] * tg[0] + dR[r * 3 + 1] * tg[1] + dR[r * 3 + 2] * tg[2] + dt[r];
      }
      if constexpr (DOF == 7)
        sc = sg + sol[6];
      cand_error = err_fn(Rc, tc, sc);
      rep.error_evals++;
      if (cand_error < curr_error)
        break;
      else if (damp < cfg->max_damp)
      {
        damp = clampd(damp * cfg->damp_inc_factor);
        solve();
      }
      else
        break;
    }
    if (cand_error >= curr_error && damp >= cfg->max_damp)
      break;
    memcpy(Rg, Rc, sizeof(Rg));
    memcpy(tg, tc, sizeof(tg));
    sg = sc;
    if (update_jac)
      prev_error = curr_error;
    curr_error = cand_error;
    damp = clampd(damp / cfg->damp_dec_factor);
    if (iter >= cfg->max_num_iters)
      break;
  }
  memcpy(R, Rg, sizeof(Rg));
  memcpy(t, tg, sizeof(tg));
  if (scale)
    *scale = sg;
  rep.iterations = (int)iter;
  rep.final_error = curr_error;
  rep.final_damp = damp;
  return rc;
}

} // namespace

extern "C" int sage_ba_track_new_frame(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *frame1,
                                       const float *code0, float scale0, const sage_ba_tracker_config *cfg, float *R, float *t,
                                       const float *match_dpts, const float *match_homo, const float *match2d, int num_matches,
                                       sage_ba_tracker_report *report)
{
  if (!ctx)
    return 1;
  try
  {
    SAGE_CHECK(kf0 && frame1 && cfg && R && t, "null argument");
    SAGE_CUDA(cudaSetDevice(ctx->device));
    const int N = kf0->N, L = kf0->L, F = kf0->F;
    const bool use_photo = cfg->use_photo != 0, use_reproj = cfg->use_reproj != 0 && num_matches > 0;
    SAGE_CHECK(use_photo || use_reproj, "no factor enabled");
    if (cfg->use_reproj && num_matches <= 3)
      throw Error{"not enough feature matches (camera_tracker.cpp:1141-1146)"};
    float *d_dpts = ctx->trk_dpts.ensure(std::max(N, 1));
    float *d_homo = ctx->trk_homo.ensure((size_t)std::max(N, 1) * 3);
    float *d_feats = ctx->trk_feats.ensure((size_t)L * std::max(N, 1) * F);
    if (use_photo)
      SAGE_CHECK(sage_ba_tracker_presample(ctx, kf0, code0, scale0, d_dpts, d_homo, d_feats) == 0, ctx->err);
    const sage_ba_camera cam = frame1->cams[0];

    sage_ba_tracker_report rep;
    memset(&rep, 0, sizeof(rep));
    auto jac_fn = [&](const float *Rg, const float *tg, float, float *AtA, float *Atb, float *err) {
      for (int i = 0; i < 36; ++i)
        AtA[i] = 0.f;
      for (int i = 0; i < 6; ++i)
        Atb[i] = 0.f;
      float e_photo = 0.f, e_rep = 0.f, A[36], b[6];
      if (use_photo)
      {
        SAGE_CHECK(sage_ba_tracker_photo_jac_error(ctx, frame1, Rg, tg, d_dpts, d_homo, d_feats, N, 0, 1.f, cfg->dpt_eps,
                                                   cfg->photo_weights, A, b, &e_photo, nullptr) == 0,
                   ctx->err);
        for (int i = 0; i < 36; ++i)
          AtA[i] += A[i];
        for (int i = 0; i < 6; ++i)
          Atb[i] += b[i];
      }
      if (use_reproj)
      {
        SAGE_CHECK(sage_ba_tracker_reproj_jac_error(ctx, &cam, Rg, tg, match_dpts, match_homo, match2d, num_matches, cfg->dpt_eps,
                                                    cfg->reproj_loss_param, cfg->reproj_weight, A, b, &e_rep, nullptr) == 0,
                   ctx->err);
        for (int i = 0; i < 36; ++i)
          AtA[i] += A[i];
        for (int i = 0; i < 6; ++i)
          Atb[i] += b[i];
      }
      *err = e_photo + e_rep;
    };
    auto err_fn = [&](const float *Rg, const float *tg, float) -> float {
      float e_photo = 0.f, e_rep = 0.f;
      if (use_photo)
        SAGE_CHECK(sage_ba_tracker_photo_error(ctx, frame1, Rg, tg, d_dpts, d_homo, d_feats, N, cfg->dpt_eps, cfg->photo_weights,
                                               &e_photo, nullptr) == 0,
                   ctx->err);
      if (use_reproj)
        SAGE_CHECK(sage_ba_tracker_reproj_error(ctx, &cam, Rg, tg, match_dpts, match_homo, match2d, num_matches, cfg->dpt_eps,
                                                cfg->reproj_loss_param, cfg->reproj_weight, &e_rep, nullptr) == 0,
                   ctx->err);
      return e_photo + e_rep;
    };
    lm_loop<6>(cfg, R, t, nullptr, -1.f, jac_fn, err_fn, rep);
    if (report)
      *report = rep;
    return 0;
  }
  catch (const sage::Error &e)
  {
    ctx->err = e.msg;
    return 1;
  }
}


// CameraTracker::TrackFrame (core/system/camera_tracker.cpp:1312-1672): the same damped Gauss-Newton loop on 7 variables,
// relative pose + depth scale of the tracked frame (UpdateVariables :467-489: scale += delta[6]; LMConvergence :527-550).
extern "C" int sage_ba_track_frame(sage_ba_context *ctx, const sage_ba_keyframe *frame0, const sage_ba_keyframe *kf1, const float *code0,
                                   const sage_ba_tracker_config *cfg, float *R, float *t, float *scale, const float *m_udpts0,
                                   const float *m_homo0, const float *m_dpts1, const float *m_homo1, int num_matches,
                                   sage_ba_tracker_report *report)
{
  if (!ctx)
    return 1;
  try
  {
    SAGE_CHECK(frame0 && kf1 && cfg && R && t && scale, "null argument");
    SAGE_CUDA(cudaSetDevice(ctx->device));
    const int N = frame0->N, L = frame0->L, F = frame0->F;
    const bool use_photo = cfg->use_photo != 0, use_mg = cfg->use_match_geom != 0;
    SAGE_CHECK(use_photo || use_mg, "at least one factor should be enabled");
    if (use_mg && num_matches <= 3)
      throw Error{"not enough feature matches (camera_tracker.cpp:1389-1393)"};
    float *d_udpts = ctx->trk_dpts.ensure(std::max(N, 1));
    float *d_homo = ctx->trk_homo.ensure((size_t)std::max(N, 1) * 3);
    float *d_feats = ctx->trk_feats.ensure((size_t)L * std::max(N, 1) * F);
    if (use_photo) // unscaled depths (dpt_map / dpt_scale) and the tracked frame's own features at its sample points
      SAGE_CHECK(sage_ba_tracker_presample(ctx, frame0, code0, 1.0f, d_udpts, d_homo, d_feats) == 0, ctx->err);

    sage_ba_tracker_report rep;
    memset(&rep, 0, sizeof(rep));
    float wsum = 0.f;
    for (int l = 0; l < L; ++l)
      wsum += cfg->photo_weights[l];
    auto jac_fn = [&](const float *Rg, const float *tg, float sg, float *AtA, float *Atb, float *err) {
      for (int i = 0; i < 49; ++i)
        AtA[i] = 0.f;
      for (int i = 0; i < 7; ++i)
        Atb[i] = 0.f;
      float e_photo = 0.f, e_mg = 0.f, A[49], b[7];
      if (use_photo)
      {
        run_tracker_photo(ctx, true, kf1, Rg, tg, d_udpts, d_homo, d_feats, N, sg, sg, cfg->dpt_eps, cfg->photo_weights, A, b, &e_photo,
                          nullptr);
        for (int i = 0; i < 49; ++i)
          AtA[i] += A[i];
        for (int i = 0; i < 7; ++i)
          Atb[i] += b[i];
      }
      if (use_mg)
      {
        run_match_geom_single(ctx, true, Rg, tg, m_udpts0, m_dpts1, m_homo0, m_homo1, num_matches, sg, sg, cfg->match_geom_loss_param,
                              cfg->match_geom_weight, A, b, &e_mg);
        for (int i = 0; i < 49; ++i)
          AtA[i] += A[i];
        for (int i = 0; i < 7; ++i)
          Atb[i] += b[i];
      }
      *err = e_photo + e_mg;
    };
    auto err_fn = [&](const float *Rg, const float *tg, float sg) -> float {
      float e_photo = 0.f, e_mg = 0.f;
      if (use_photo)
        run_tracker_photo(ctx, false, kf1, Rg, tg, d_udpts, d_homo, d_feats, N, sg, 0.f, cfg->dpt_eps, cfg->photo_weights, nullptr, nullptr,
                          &e_photo, nullptr);
      if (use_mg)
        run_match_geom_single(ctx, false, Rg, tg, m_udpts0, m_dpts1, m_homo0, m_homo1, num_matches, sg, 0.f, cfg->match_geom_loss_param,
                              cfg->match_geom_weight, nullptr, nullptr, &e_mg);
      return e_photo + e_mg;
    };
    const int rc = lm_loop<7>(cfg, R, t, scale, use_mg ? -1.f : wsum * 9.9f, jac_fn, err_fn, rep); // no overlap: :1500-1504
    if (report)
      *report = rep;
    return rc;
  }
  catch (const sage::Error &e)
  {
    ctx->err = e.msg;
    return 1;
  }
}

extern "C" {

// The LM loop of sage_ba_track_new_frame (dof 6) / sage_ba_track_frame (dof 7) over a caller-supplied cost.  Host code only.
int sage_ba_tracker_lm_callbacks(int dof, const sage_ba_tracker_config *cfg, float *R, float *t, float *scale, sage_ba_lm_jac_callback jac,
                                 sage_ba_lm_err_callback err, void *user, sage_ba_tracker_report *report)
{
  if (!cfg || !R || !t || !jac || !err || (dof != 6 && dof != 7) || (dof == 7 && !scale))
    return 1;
  sage_ba_tracker_report rep;
  memset(&rep, 0, sizeof(rep));
  auto jac_fn = [&](const float *Rg, const float *tg, float sg, float *AtA, float *Atb, float *e) { jac(user, Rg, tg, sg, AtA, Atb, e); };
  auto err_fn = [&](const float *Rg, const float *tg, float sg) -> float { return err(user, Rg, tg, sg); };
  const int rc = dof == 6 ? lm_loop<6>(cfg, R, t, nullptr, -1.f, jac_fn, err_fn, rep) : lm_loop<7>(cfg, R, t, scale, -1.f, jac_fn, err_fn, rep);
  if (report)
    *report = rep;
  return rc;
}

/* The tracker's linear solve, exposed for callers that keep their own LM loop and for the parity tests:
 * x = (AtA + damp * diag(AtA)).colPivHouseholderQr().solve(Atb), n = 6 | 7, float (camera_tracker.cpp:1182-1183). */
int sage_ba_tracker_solve(const float *AtA, const float *Atb, int n, float damp, float *x)
{
  if (!AtA || !Atb || !x || n < 1 || n > 8)
    return 1;
  float damped[64];
  for (int i = 0; i < n * n; ++i)
    damped[i] = AtA[i];
  for (int i = 0; i < n; ++i)
    damped[i * n + i] = AtA[i * n + i] + damp * AtA[i * n + i];
  return solve_small(damped, Atb, n, x) ? 0 : 1;
}

/* se3_exp<float> (core/mapping/mapping_utils.h:316-346) as the tracker's update uses it: R [9] row-major, t [3]. */
int sage_ba_se3_exp(const float *omega, const float *v, float *R, float *t)
{
  if (!omega || !v || !R || !t)
    return 1;
  se3_exp_host(omega, v, R, t);
  return 0;
}

} // extern "C"

// problem.cu -- batched local bundle adjustment: device-resident state, one launch per factor kind, block-sparse fp64
// normal equations assembled in a fixed order, elimination by the hand-written block Cholesky of blocksolve.cu (per-keyframe
// Schur complements), keyframe-owner sharding with the collective issued from C++ (NCCL), and the host LM loop.  New design: the reference hands this to GTSAM ISAM2
// (core/mapping/mapper.cpp:544); the per-factor arithmetic is the reference's (a1, a4, a5 of SURVEY.md
// section 8), the priors are CodeFactor / ScaleFactor (core/gtsam/code_factor.cpp:42-104,
// scale_factor.cpp:115-130) and the retraction is the left-multiplicative one of
// core/gtsam/gtsam_traits.h:45-70.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cstdlib>

#include "blocksolve.h"
#include "comm.h"
#include "sage_internal.h"

using namespace sage;

namespace sage
{

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rel_pose(const float *P0, const float *P1, float *R10, float *t10)
{
  // R10 = R1^T R0, t10 = R1^T (t0 - t1)   (core/gtsam/photometric_factor.cpp:280-281), fp32
  const float *R0 = P0, *t0 = P0 + 9, *R1 = P1, *t1 = P1 + 9;
  for (int r = 0; r < 3; ++r)
  {
    for (int c = 0; c < 3; ++c)
      R10[r * 3 + c] = R1[0 * 3 + r] * R0[0 * 3 + c] + R1[1 * 3 + r] * R0[1 * 3 + c] + R1[2 * 3 + r] * R0[2 * 3 + c];
    t10[r] = R1[0 * 3 + r] * (t0[0] - t1[0]) + R1[1 * 3 + r] * (t0[1] - t1[1]) + R1[2 * 3 + r] * (t0[2] - t1[2]);
  }
}

__global__ void setup_photo_kernel(PhotoFactor *f, const int2 *ij, const int2 *offs, int n, const float *poses, const float *codes,
                                   const float *scales, int C, float eps, int jac)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n)
    return;
  const int i = ij[t].x, j = ij[t].y;
  const float *P0 = poses + i * 12, *P1 = poses + j * 12;
  PhotoFactor &o = f[t];
  rel_pose(P0, P1, o.R10, o.t10);
  for (int k = 0; k < 9; ++k)
  {
    o.R0[k] = P0[k];
    o.R1[k] = P1[k];
  }
  for (int k = 0; k < 3; ++k)
    o.t0[k] = P0[9 + k];
  for (int k = 0; k < C; ++k)
    o.code0[k] = codes[i * C + k];
  o.scale0 = scales[i];
  o.eps = eps;
  o.out = jac ? offs[t].x : offs[t].y;
}

__global__ void setup_geo_kernel(GeoFactor *f, const int2 *ij, const int2 *offs, int n, const float *poses, const float *codes,
                                 const float *scales, int C, float eps, int jac)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n)
    return;
  const int i = ij[t].x, j = ij[t].y;
  const float *P0 = poses + i * 12, *P1 = poses + j * 12;
  GeoFactor &o = f[t];
  rel_pose(P0, P1, o.R10, o.t10);
  for (int k = 0; k < 9; ++k)
  {
    o.R0[k] = P0[k];
    o.R1[k] = P1[k];
  }
  for (int k = 0; k < 3; ++k)
    o.t0[k] = P0[9 + k];
  for (int k = 0; k < C; ++k)
    o.code0[k] = codes[i * C + k];
  o.scale0 = scales[i];
  o.scale1 = scales[j];
  o.dscale = scales[j];
  o.eps = eps;
  o.out = jac ? offs[t].x : offs[t].y;
}

__global__ void setup_reproj_kernel(ReprojFactor *f, const int2 *ij, const int2 *offs, int n, const float *poses, const float *codes,
                                    const float *scales, int C, float eps, int jac)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n)
    return;
  const int i = ij[t].x, j = ij[t].y;
  const float *P0 = poses + i * 12, *P1 = poses + j * 12;
  ReprojFactor &o = f[t];
  rel_pose(P0, P1, o.R10, o.t10);
  for (int k = 0; k < 9; ++k)
  {
    o.R0[k] = P0[k];
    o.R1[k] = P1[k];
  }
  for (int k = 0; k < 3; ++k)
    o.t0[k] = P0[9 + k];
  for (int k = 0; k < C; ++k)
    o.code0[k] = codes[i * C + k];
  o.scale0 = scales[i];
  o.eps = eps;
  o.out = jac ? offs[t].x : offs[t].y;
}

// D[k][p] = bias_k[p] + basis_k[p,:] . code_k   for every keyframe k (grid.y)
struct KfMaps
{
  const float *bias, *basis, *mask;
  float4 *dgm;
  float *dscr;
  int k; // keyframe index (row of the state arrays)
};

__global__ void depth_unscaled_batched_kernel(const KfMaps *maps, const float *codes, int HW, int C)
{
  __shared__ float sc[SAGE_MAX_CODE];
  const int k = maps[blockIdx.y].k;
  if (threadIdx.x < C)
    sc[threadIdx.x] = codes[k * C + threadIdx.x];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW)
    return;
  const KfMaps m = maps[blockIdx.y];
  float acc = 0.f;
  const float4 *row = reinterpret_cast<const float4 *>(m.basis + (size_t)p * C);
  for (int q = 0; q < C / 4; ++q)
  {
    const float4 v = __ldg(row + q);
    acc += v.x * sc[4 * q] + v.y * sc[4 * q + 1] + v.z * sc[4 * q + 2] + v.w * sc[4 * q + 3];
  }
  m.dscr[p] = __ldg(m.bias + p) + acc;
}

// UpdateDepth for every keyframe (core/mapping/mapping_utils.h:216-222): dpt_map = scale * (bias + basis . code), in the reference's
// operation order (GEMV first, then the bias, then the scale).  out: [K][HW]
__global__ void update_depth_batched_kernel(const KfMaps *maps, const float *codes, const float *scales, int HW, int C, float *out)
{
  __shared__ float sc[SAGE_MAX_CODE];
  const int k = maps[blockIdx.y].k;
  if (threadIdx.x < C)
    sc[threadIdx.x] = codes[k * C + threadIdx.x];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW)
    return;
  const KfMaps m = maps[blockIdx.y];
  float acc = 0.f;
  const float4 *row = reinterpret_cast<const float4 *>(m.basis + (size_t)p * C);
  for (int c = 0; c < C / 4; ++c)
  {
    const float4 b = __ldg(row + c);
    acc = fmaf(b.x, sc[4 * c + 0], acc);
    acc = fmaf(b.y, sc[4 * c + 1], acc);
    acc = fmaf(b.z, sc[4 * c + 2], acc);
    acc = fmaf(b.w, sc[4 * c + 3], acc);
  }
  out[(size_t)k * HW + p] = scales[k] * (__ldg(m.bias + p) + acc);
}

__global__ void depth_pack_batched_kernel(const KfMaps *maps, int H, int W)
{
  const KfMaps m = maps[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W)
    return;
  const int y = p / W, x = p - y * W;
  const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
  const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
  const float *D = m.dscr;
  m.dgm[p] = make_float4(D[p], 0.5f * (D[y * W + xp] - D[y * W + xm]), 0.5f * (D[yp * W + x] - D[ym * W + x]), m.mask[p]);
}

// priors: CodeFactor (AtA = w I, Atb = w (init - code), err = w mean((init-code)^2)) and ScaleFactor
// (AtA = w/s^2, Atb = w/s (log s0 - log s), err = w (log s0 - log s)^2).  One block.
__global__ void priors_kernel(const PriorSpec *pr, int np, const float *codes, const float *scales, double *H, double *g,
                              double *prior_cost, int n, int K, int C, int add_to_system)
{
  __shared__ double red[256];
  double cost = 0.0;
  // one (prior, component) pair per thread step; priors on the same variable may coexist, hence the atomics
  for (int e = threadIdx.x; e < np * C; e += blockDim.x)
  {
    const int q = e / C, c = e - q * C;
    const PriorSpec &p = pr[q];
    const int cb = 6 * K + p.kf * (C + 1);
    if (p.kind == 0)
    {
      const double diff = (double)p.init_code[c] - (double)codes[p.kf * C + c];
      cost += (double)p.weight * diff * diff / (double)C;
      if (add_to_system)
      {
        atomicAdd(&H[(size_t)(cb + c) * n + cb + c], (double)p.weight);
        atomicAdd(&g[cb + c], (double)p.weight * diff);
      }
    }
    else if (c == 0)
    {
      const double s = (double)scales[p.kf];
      const double d = log((double)p.init_scale) - log(s);
      cost += (double)p.weight * d * d;
      if (add_to_system)
      {
        atomicAdd(&H[(size_t)(cb + C) * n + cb + C], (double)p.weight / (s * s));
        atomicAdd(&g[cb + C], (double)p.weight / s * d);
      }
    }
  }
  red[threadIdx.x] = cost;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1)
  {
    if (threadIdx.x < s)
      red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *prior_cost = red[0];
}

// total = sum_f buf[pos_f] + prior   (fixed order, fp64)
__global__ void total_cost_kernel(const float *buf, const int *pos, int nf, const double *prior, double *total)
{
  __shared__ double red[256];
  double s = 0.0;
  for (int f = threadIdx.x; f < nf; f += blockDim.x)
    s += (double)buf[pos[f]];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = blockDim.x / 2; k > 0; k >>= 1)
  {
    if (threadIdx.x < k)
      red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *total = red[0] + *prior;
}

// Dense cross-check solver: Hd = H with damped diagonal (fixed variables become identity rows / columns with zero gradient),
// written with the code+scale block FIRST and the pose block LAST (variable v moves to
// v + nc for poses, v - np for the rest), so that one Cholesky factorisation of the whole matrix eliminates the code block,
// forms the Schur complement onto the pose block in its trailing sub-matrix and factors it (block Cholesky == Schur).
__global__ void damp_perm_kernel(const double *H, const double *g, const unsigned char *fixed, double *Hd, double *gd, int n, int np,
                                 double damp)
{
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n * n)
    return;
  const int r = (int)(e / n), c = (int)(e % n);
  const int nc = n - np;
  const int pr = r < np ? r + nc : r - np, pc = c < np ? c + nc : c - np;
  double v = H[e];
  if (fixed[r] || fixed[c])
    v = (r == c) ? 1.0 : 0.0;
  else if (r == c)
  {
    v = v + damp * v;
    if (!(v > 0.0))
      v = 1.0;
  }
  Hd[(size_t)pr * n + pc] = v;
  if (c == 0)
    gd[pr] = fixed[r] ? 0.0 : g[r];
}

__global__ void unpermute_kernel(const double *in, double *out, int n, int np)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n)
    out[v] = in[v < np ? v + (n - np) : v - np];
}

__device__ void se3_exp_dev(const float *w, const float *v, float *R, float *t)
{
  // se3_exp (core/mapping/mapping_utils.h:316-346), fp32
  float theta = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  float nx = 1.f, ny = 0.f, nz = 0.f;
  if (theta > 0.f)
  {
    nx = w[0] / theta;
    ny = w[1] / theta;
    nz = w[2] / theta;
  }
  theta = fmaxf(theta, 1.0e-14f);
  const float s = sinf(theta), c = cosf(theta);
  const float K[9] = {0.f, -nz, ny, nz, 0.f, -nx, -ny, nx, 0.f};
  float K2[9];
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q)
      K2[r * 3 + q] = K[r * 3 + 0] * K[0 * 3 + q] + K[r * 3 + 1] * K[1 * 3 + q] + K[r * 3 + 2] * K[2 * 3 + q];
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

// candidate = x (+) delta : pose <- exp([v, w]) * pose (left multiplication), code += dc, scale += ds
__global__ void retract_kernel(const float *poses, const float *codes, const float *scales, const double *delta, float *poses_c,
                               float *codes_c, float *scales_c, int K, int C)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K)
    return;
  float v[3], w[3], dR[9], dt[3];
  for (int q = 0; q < 3; ++q)
  {
    v[q] = (float)delta[6 * k + q];
    w[q] = (float)delta[6 * k + 3 + q];
  }
  se3_exp_dev(w, v, dR, dt);
  const float *R = poses + k * 12, *t = R + 9;
  float *Rc = poses_c + k * 12, *tc = Rc + 9;
  for (int r = 0; r < 3; ++r)
  {
    for (int q = 0; q < 3; ++q)
      Rc[r * 3 + q] = dR[r * 3 + 0] * R[0 * 3 + q] + dR[r * 3 + 1] * R[1 * 3 + q] + dR[r * 3 + 2] * R[2 * 3 + q];
    tc[r] = dR[r * 3 + 0] * t[0] + dR[r * 3 + 1] * t[1] + dR[r * 3 + 2] * t[2] + dt[r];
  }
  const int cb = 6 * K + k * (C + 1);
  for (int c = 0; c < C; ++c)
    codes_c[k * C + c] = codes[k * C + c] + (float)delta[cb + c];
  scales_c[k] = scales[k] + (float)delta[cb + C];
}

} // namespace sage


struct sage_ba_problem
{
  sage_ba_context *ctx = nullptr;
  int K = 0, F = 0, C = 0, L = 0, H = 0, W = 0, N = 0;
  std::vector<sage_ba_keyframe *> kfs; // entries of keyframes this rank never touches may be null
  sage_ba_keyframe *kf_any = nullptr;  // any non-null keyframe (shapes, cameras)
  float eps = 1e-4f;
  int rank = 0, world = 1;
  bool built = false;

  // factor specs in order of addition (global index = position in `metas`); typed lists hold only this rank's factors
  std::vector<FactorMeta> metas;
  std::vector<PhotoFactor> photo_h;
  std::vector<GeoFactor> geo_h;
  std::vector<ReprojFactor> reproj_h;
  std::vector<int> photo_g, geo_g, reproj_g; // global factor index of each typed factor
  // per-factor payload kept on the host until the problem is built (ownership -- and with it which rank needs the device
  // data -- is only known once every factor has been added)
  struct FactorSpec
  {
    std::vector<float> w;      // photometric level weights
    float loss = 0.f, weight = 0.f;
    std::vector<int32_t> loc;  // reprojection matches
    std::vector<float> homo, uv;
  };
  std::vector<FactorSpec> specs;
  std::vector<PriorSpec> priors;
  std::vector<unsigned char> fixed_h;
  std::vector<void *> owned; // device allocations owned by the problem (match arrays)
  size_t fbuf_count = 0, fseg = 0, cseg = 0; // packed buffers: `world` segments of fseg / cseg floats, one per owner
  long residuals = 0;

  // device
  DevBuf<PhotoFactor> photo_d;
  DevBuf<GeoFactor> geo_d;
  DevBuf<ReprojFactor> reproj_d;
  DevBuf<int2> photo_ij, geo_ij, reproj_ij, photo_off, geo_off, reproj_off;
  int n_photo = 0, n_geo = 0, n_reproj = 0; // this shard
  DevBuf<FactorMeta> metas_d;
  DevBuf<PriorSpec> priors_d;
  DevBuf<int> errpos_d, costpos_d;
  DevBuf<KfMaps> maps_d, maps_all_d;
  int n_maps = 0; // keyframes whose depth map this rank rebuilds per state (targets of its geometric factors)
  DevBuf<float> state[2][3]; // [which][poses, codes, scales]
  DevBuf<float> fbuf, cbuf, partH, partE, partHg, partEg, depth_out; // per-CTA partials: photometric / geometric (they run concurrently)
  DevBuf<double> Hm, gv, Hd, gd, delta, prior_cost, total_cost, work; // Hm .. gd / work: dense cross-check path only
  DevBuf<int> info;
  DevBuf<unsigned char> fixed_d;
  PinBuf<double> hcost;
  PinBuf<int> hinfo;
  int potrf_lwork = 0;
  int solver = 0;      // 0: block Cholesky, nested-dissection order; 1: dense fused-Schur potrf (cuSOLVER, cross-check); 2: block Cholesky, natural order
  bool staged = false; // opt-in (SAGE_BA_STAGED=1): photometric kernels read levels >= 1 from TMA-staged windows; measured slower, see profiles/README.md
  bool dense_valid = false;
  bool lin_valid = false;         // fbuf / H hold the linearisation at state[0]
  bool relinearize_always = false; // lm_step: linearise even when the state did not change since the last linearisation
  bool deterministic = false;      // CTA decomposition independent of the rank count (bit-identical results for any world size)
  bool concurrent_factors = true;  // photometric / geometric / reprojection launches on forked streams
  bool geo_tc = false;             // the geometric lineariser this problem was built for (snapshot of the process-wide switch)
  BlockSystem bs;
  int slices_photo = 32, slices_geo = 32, slices_photo_err = 32, slices_geo_err = 32; // CTAs per factor (linearise / error-only)

  sage_ba_allreduce_fn allreduce = nullptr;
  void *allreduce_user = nullptr;
  sage_ba_comm *comm = nullptr;

  // optional CUDA-event profiling of the launches (bench.py's roofline leg)
  bool profiling = false;
  struct ProfSpan
  {
    int kind;
    cudaEvent_t a, b;
  };
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[SAGE_BA_PROF_KINDS] = {0};
  long prof_n[SAGE_BA_PROF_KINDS] = {0};

  int dim() const { return K * (7 + C); }
};

namespace sage
{

struct ProfScope
{
  sage_ba_problem *p;
  cudaEvent_t b = nullptr;
  cudaStream_t st;
  ProfScope(sage_ba_problem *p_, int kind, cudaStream_t stream = nullptr) : p(p_), st(stream ? stream : p_->ctx->stream)
  {
    if (!p->profiling)
      return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!p->event_pool.empty())
      {
        e = p->event_pool.back();
        p->event_pool.pop_back();
      }
      else
        cudaEventCreate(&e);
      return e;
    };
    cudaEvent_t a = get();
    b = get();
    cudaEventRecord(a, st);
    p->spans.push_back({kind, a, b});
  }
  void end()
  {
    if (b)
      cudaEventRecord(b, st);
    b = nullptr;
  }
  ~ProfScope() { end(); }
};

static const sage_ba_keyframe *need_kf(sage_ba_problem *p, int k, bool depth, bool feats, bool samples)
{
  const sage_ba_keyframe *kf = p->kfs[k];
  SAGE_CHECK(kf, "a keyframe this rank's factors touch was passed as null");
  SAGE_CHECK(!depth || (kf->bias && kf->basis), "keyframe lacks depth data");
  SAGE_CHECK(!feats || (kf->fg && kf->mask), "keyframe lacks feature maps / mask");
  SAGE_CHECK(!samples || (kf->loc1d && kf->homo && kf->N == p->N), "keyframe lacks sample data (or sample counts differ)");
  return kf;
}

static void problem_build(sage_ba_problem *p)
{
  if (p->built)
    return;
  sage_ba_context *ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  const int K = p->K, C = p->C;
  // ---- ownership: the distinct ordered pairs, sorted by (host keyframe, target), are cut into `world` equal runs; every factor
  // of a pair goes to the pair's rank.  A keyframe's pairs are consecutive in that order, so its maps are read by one rank
  // (two at a cut), and the runs differ by at most one pair however uneven the keyframes' degrees are.
  {
    std::vector<int> pi(p->metas.size()), pj(p->metas.size()), own(p->metas.size());
    for (size_t f = 0; f < p->metas.size(); ++f)
    {
      pi[f] = p->metas[f].i;
      pj[f] = p->metas[f].j;
    }
    sage_ba_shard_plan((int)p->metas.size(), pi.data(), pj.data(), p->world, own.data());
    for (size_t f = 0; f < p->metas.size(); ++f)
      p->metas[f].owner = own[f];
  }
  // ---- typed device factors of the pairs this rank owns
  for (size_t f = 0; f < p->metas.size(); ++f)
  {
    const FactorMeta &m = p->metas[f];
    if (m.owner != p->rank)
      continue;
    const sage_ba_problem::FactorSpec &sp = p->specs[f];
    const int i = m.i, j = m.j;
    if (m.kind == 0)
    {
      const sage_ba_keyframe *a = need_kf(p, i, true, true, true), *b = need_kf(p, j, false, true, false);
      SAGE_CHECK(a->sfeat, "keyframe lacks pre-sampled features");
      PhotoFactor pf;
      memset(&pf, 0, sizeof(pf));
      pf.fg0 = a->fg;
      pf.fg1 = b->fg;
      pf.mask1 = b->mask;
      pf.bias0 = a->bias;
      pf.basis0 = a->basis;
      if (p->staged)
      {
        ensure_sorted_samples(ctx, p->kfs[i]);
        pf.sfeat0 = a->sfeat_s;
        pf.loc1d = a->loc1d_s;
        pf.homo = a->homo_s;
      }
      else
      {
        pf.sfeat0 = a->sfeat;
        pf.loc1d = a->loc1d;
        pf.homo = a->homo;
      }
      pf.N = a->N;
      memcpy(pf.w, sp.w.data(), sizeof(float) * p->L);
      p->photo_h.push_back(pf);
      p->photo_g.push_back((int)f);
    }
    else if (m.kind == 1)
    {
      const sage_ba_keyframe *a = need_kf(p, i, true, false, true), *b = need_kf(p, j, true, false, false);
      SAGE_CHECK(b->mask && b->dgm, "keyframe lacks mask / depth-map buffers");
      GeoFactor gf;
      memset(&gf, 0, sizeof(gf));
      gf.bias0 = a->bias;
      gf.basis0 = a->basis;
      gf.loc1d = a->loc1d;
      gf.homo = a->homo;
      gf.dgm1 = b->dgm;
      gf.basis1 = b->basis;
      gf.N = a->N;
      gf.loss_param = sp.loss;
      gf.weight = sp.weight;
      p->geo_h.push_back(gf);
      p->geo_g.push_back((int)f);
    }
    else
    {
      const sage_ba_keyframe *a = need_kf(p, i, true, false, false);
      const int M = (int)sp.loc.size();
      float *dm = nullptr;
      SAGE_CUDA(cudaMalloc(&dm, sizeof(float) * 6 * M));
      p->owned.push_back(dm);
      std::vector<float> hm((size_t)6 * M);
      memcpy(hm.data(), sp.homo.data(), sizeof(float) * 3 * M);
      memcpy(hm.data() + 3 * M, sp.uv.data(), sizeof(float) * 2 * M);
      memcpy(hm.data() + 5 * M, sp.loc.data(), sizeof(int32_t) * M);
      SAGE_CUDA(cudaMemcpy(dm, hm.data(), sizeof(float) * 6 * M, cudaMemcpyHostToDevice));
      ReprojFactor rf;
      memset(&rf, 0, sizeof(rf));
      rf.bias0 = a->bias;
      rf.basis0 = a->basis;
      rf.homo = dm;
      rf.match2d = dm + 3 * M;
      rf.loc1d = reinterpret_cast<const int *>(dm + 5 * M);
      rf.M = M;
      rf.loss_param = sp.loss;
      rf.weight = sp.weight;
      rf.fx = a->cams[0].fx;
      rf.fy = a->cams[0].fy;
      rf.cx = a->cams[0].u0;
      rf.cy = a->cams[0].v0;
      p->reproj_h.push_back(rf);
      p->reproj_g.push_back((int)f);
    }
  }
  // ---- packed output buffers: one segment per owner rank (all-gather friendly), factors in order of addition inside it
  {
    std::vector<size_t> used(p->world, 0), cnt(p->world, 0);
    for (FactorMeta &m : p->metas)
    {
      m.off = (int)used[m.owner];
      m.cost_off = (int)(2 * cnt[m.owner]);
      used[m.owner] += (size_t)m.D * m.D + m.D + 2;
      cnt[m.owner] += 1;
    }
    size_t fmax = 4, cmax = 4;
    for (int r = 0; r < p->world; ++r)
    {
      fmax = std::max(fmax, used[r]);
      cmax = std::max(cmax, 2 * cnt[r]);
    }
    p->fseg = (fmax + 31) / 32 * 32;
    p->cseg = (cmax + 31) / 32 * 32;
    for (FactorMeta &m : p->metas)
    {
      m.off += (int)(m.owner * p->fseg);
      m.cost_off += (int)(m.owner * p->cseg);
    }
    p->fbuf_count = p->fseg * p->world;
  }
  std::vector<int2> pij, gij, rij, poff, goff, roff;
  for (size_t q = 0; q < p->photo_h.size(); ++q)
  {
    const FactorMeta &m = p->metas[p->photo_g[q]];
    pij.push_back(make_int2(m.i, m.j));
    poff.push_back(make_int2(m.off, m.cost_off));
  }
  for (size_t q = 0; q < p->geo_h.size(); ++q)
  {
    const FactorMeta &m = p->metas[p->geo_g[q]];
    gij.push_back(make_int2(m.i, m.j));
    goff.push_back(make_int2(m.off, m.cost_off));
  }
  for (size_t q = 0; q < p->reproj_h.size(); ++q)
  {
    const FactorMeta &m = p->metas[p->reproj_g[q]];
    rij.push_back(make_int2(m.i, m.j));
    roff.push_back(make_int2(m.off, m.cost_off));
  }
  p->n_photo = (int)p->photo_h.size();
  p->n_geo = (int)p->geo_h.size();
  p->n_reproj = (int)p->reproj_h.size();
  auto up = [&](auto &dev, const auto &host) {
    using T = typename std::remove_reference<decltype(host[0])>::type;
    if (host.empty())
      return;
    auto *d = dev.ensure(host.size());
    SAGE_CUDA(cudaMemcpyAsync((void *)d, (const void *)host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  };
  up(p->photo_d, p->photo_h);
  up(p->geo_d, p->geo_h);
  up(p->reproj_d, p->reproj_h);
  up(p->photo_ij, pij);
  up(p->geo_ij, gij);
  up(p->reproj_ij, rij);
  up(p->photo_off, poff);
  up(p->geo_off, goff);
  up(p->reproj_off, roff);
  up(p->metas_d, p->metas);
  up(p->priors_d, p->priors);
  std::vector<int> errpos, costpos;
  for (const FactorMeta &m : p->metas)
  {
    errpos.push_back(m.off + m.D * m.D + m.D);
    costpos.push_back(m.cost_off);
  }
  up(p->errpos_d, errpos);
  up(p->costpos_d, costpos);
  // depth maps are rebuilt per state only for the targets of this rank's geometric factors
  std::vector<KfMaps> maps, maps_all;
  {
    std::vector<char> need(K, 0);
    for (size_t q = 0; q < p->geo_h.size(); ++q)
      need[p->metas[p->geo_g[q]].j] = 1;
    for (int k = 0; k < K; ++k)
    {
      sage_ba_keyframe *kf = p->kfs[k];
      if (kf && kf->bias)
        maps_all.push_back(KfMaps{kf->bias, kf->basis, kf->mask, kf->dgm, kf->dscr, k});
      if (need[k])
        maps.push_back(KfMaps{kf->bias, kf->basis, kf->mask, kf->dgm, kf->dscr, k});
    }
  }
  p->n_maps = (int)maps.size();
  up(p->maps_d, maps);
  up(p->maps_all_d, maps_all);
  if (p->fixed_h.empty())
    p->fixed_h.assign(p->dim(), 0);
  up(p->fixed_d, p->fixed_h);
  p->bs.build(K, C, p->metas, p->priors, p->fixed_h, p->solver == 2 ? 1 : (getenv("SAGE_BA_NATURAL_ORDER") ? 1 : 0), s);

  p->fbuf.ensure(std::max<size_t>(p->fbuf_count, 4));
  p->cbuf.ensure(std::max<size_t>(p->cseg * p->world, 4));
  SAGE_CUDA(cudaMemsetAsync(p->fbuf.p, 0, p->fbuf.cap * sizeof(float), s));
  SAGE_CUDA(cudaMemsetAsync(p->cbuf.p, 0, p->cbuf.cap * sizeof(float), s));
  // slices = CTAs per factor.  Two policies:
  //  * default: end the launch on a full wave -- slices = floor(waves * resident CTAs / owned factors) for the largest wave count
  //    that still leaves every CTA a couple of thousand samples.  The factor kernels run only a handful of waves and a CTA costs
  //    ~20 us beyond its samples (its four warps drift apart, the slot idles until the last one is done), so coarse slices that
  //    fill whole waves are worth 10 % (32 KF / 180 pairs, B200: lineariser 6.97 ms at 12 slices, 7.7 ms at 40, 8.0 ms at 80).
  //    The count then depends on how many factors the rank owns, i.e. on the number of GPUs: a factor's partial sums are added
  //    in a different order, its outputs agree across GPU counts to fp32 round-off (1e-7), not bit for bit.
  //  * deterministic (sage_ba_problem_set_deterministic, SAGE_BA_DETERMINISTIC=1): the count depends on the problem only (samples
  //    per keyframe).  With the fixed-order assembly and the flag-ordered solver the whole LM trajectory is then BIT-identical
  //    for every number of GPUs (tests/test_gpu_multirank.py), at the price above.
  auto per_cta = [&](const char *env, int dflt) {
    const char *e = getenv(env);
    const int spc = e ? std::max(32, atoi(e)) : dflt;
    return std::max(1, (p->N + spc - 1) / spc);
  };
  auto pick = [&](int nfac, int ctas_per_sm, int max_waves, int min_samples_per_cta) {
    if (nfac <= 0)
      return 1;
    const int slots = std::max(1, ctas_per_sm) * ctx->num_sms;
    const int cap = std::max(1, p->N / std::max(1, min_samples_per_cta));
    for (int w = max_waves; w >= 1; --w)
    {
      const int sl = w * slots / nfac;
      if (sl >= 1 && sl <= cap)
        return sl;
    }
    return std::max(1, std::min(cap, slots / nfac));
  };
  p->geo_tc = geo_uses_tc(true, C);
  if (p->deterministic)
  {
    p->slices_photo = per_cta("SAGE_BA_SPC_PHOTO", 2048);
    p->slices_photo_err = per_cta("SAGE_BA_SPC_PHOTO_ERR", 2048);
    p->slices_geo = per_cta("SAGE_BA_SPC_GEO", 2048);
    p->slices_geo_err = per_cta("SAGE_BA_SPC_GEO_ERR", 1024);
  }
  else
  {
    p->slices_photo = pick(p->n_photo, photo_ctas_per_sm(PH_MAP_JAC, p->F, C, p->staged), 5, 2048);
    p->slices_photo_err = pick(p->n_photo, photo_ctas_per_sm(PH_MAP_ERR, p->F, C, p->staged), 5, 2048);
    p->slices_geo = pick(p->n_geo, geo_ctas_per_sm(true, C, p->geo_tc), 7, 1024);
    if (p->geo_tc && p->n_geo > 0)
    {
      // The tcgen05 lineariser's CTAs cost a fixed amount to start and to finish (tensor-memory allocation, a cold first round, a
      // 57 KB partial the finalize kernel reads back) and its time follows the wave efficiency (sweeps of 3..23 slices at 1 and 2
      // GPUs, 67 slices of 9 rounds at 8 GPUs: +60 %).  Choose the slice count that maximises
      // (filled fraction of the last wave) x (useful fraction of a CTA's life), with at least 32 rounds per CTA.
      const int slots = std::max(1, geo_ctas_per_sm(true, C, true)) * ctx->num_sms;
      const int cap = std::max(1, p->N / 4096);
      double best = -1.0;
      for (int w = 1; w <= 7; ++w)
      {
        const int sl = std::min(cap, std::max(1, w * slots / p->n_geo));
        const double waves = (double)sl * p->n_geo / slots, rounds = (double)p->N / sl / 128.0;
        const double score = waves / std::ceil(waves) * rounds / (rounds + 1.5);
        if (score > best + 1e-9)
          best = score, p->slices_geo = sl;
      }
    }
    p->slices_geo_err = pick(p->n_geo, geo_ctas_per_sm(false, C, false), 8, 512);
    if (const char *e = getenv("SAGE_BA_SLICES_PHOTO"))
      p->slices_photo = std::max(1, atoi(e));
    if (const char *e = getenv("SAGE_BA_SLICES_GEO"))
      p->slices_geo = std::max(1, atoi(e));
    if (const char *e = getenv("SAGE_BA_SLICES_PHOTO_ERR"))
      p->slices_photo_err = std::max(1, atoi(e));
    if (const char *e = getenv("SAGE_BA_SLICES_GEO_ERR"))
      p->slices_geo_err = std::max(1, atoi(e));
  }
  if (getenv("SAGE_BA_DEBUG"))
    fprintf(stderr, "[sage_ba] slices: photo %d / %d (err), geo %d / %d (err); CTAs per SM: photo %d, geo %d (tcgen05 %d)\n", p->slices_photo,
            p->slices_photo_err, p->slices_geo, p->slices_geo_err, photo_ctas_per_sm(PH_MAP_JAC, p->F, C, p->staged), geo_ctas_per_sm(true, C, p->geo_tc),
            (int)p->geo_tc);
  p->partH.ensure(std::max<size_t>((size_t)p->n_photo * p->slices_photo * photo_partial_floats(PH_MAP_JAC, C), 4));
  p->partHg.ensure(std::max<size_t>((size_t)p->n_geo * p->slices_geo * geo_partial_floats(true, C, p->geo_tc), 4));
  p->partE.ensure(std::max<size_t>(2 * (size_t)p->n_photo * std::max(p->slices_photo, p->slices_photo_err), 4));
  p->partEg.ensure(std::max<size_t>(2 * (size_t)p->n_geo * std::max(p->slices_geo, p->slices_geo_err), 4));
  const int n = p->dim();
  p->delta.ensure(n);
  SAGE_CUDA(cudaMemsetAsync(p->delta.p, 0, sizeof(double) * n, s));
  p->prior_cost.ensure(2);
  p->total_cost.ensure(2);
  p->info.ensure(4);
  p->hcost.ensure(4);
  p->hinfo.ensure(4);
  SAGE_CUDA(cudaStreamSynchronize(s));
  p->built = true;
}

// dense copies of H / g (tests, the cuSOLVER cross-check solver); allocated on first use
static void ensure_dense(sage_ba_problem *p)
{
  sage_ba_context *ctx = p->ctx;
  const int n = p->dim();
  p->Hm.ensure((size_t)n * n);
  p->gv.ensure(n);
  if (!p->dense_valid)
  {
    p->bs.expand_dense(p->Hm.p, p->gv.p, n, ctx->stream, &ctx->launches);
    p->dense_valid = true;
  }
}

// Linearise (jac) or evaluate this rank's factors at state `which` into `out`.  The three factor kinds are independent: the
// photometric launches stay on the main stream, the geometric chain (per-state depth maps -> geometric kernel) and the
// reprojection kernel run on two side streams forked off it and joined at the end, so the small launches and the tails of the
// big ones overlap instead of queueing (worth ~0.3 ms per iteration at 32 keyframes, more when a rank owns few pairs).
static void run_factors(sage_ba_problem *p, int which, bool jac, float *out)
{
  sage_ba_context *ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  if (!ctx->aux[0])
  {
    for (int q = 0; q < 2; ++q)
    {
      SAGE_CUDA(cudaStreamCreateWithFlags(&ctx->aux[q], cudaStreamNonBlocking));
      SAGE_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[q], cudaEventDisableTiming));
    }
    SAGE_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  }
  const bool serial = !p->concurrent_factors || p->profiling; // per-kind event timing needs the kinds one after the other
  cudaStream_t sg = serial ? s : ctx->aux[0], sr = serial ? s : ctx->aux[1];
  const float *poses = p->state[which][0].p, *codes = p->state[which][1].p, *scales = p->state[which][2].p;
  const sage_ba_keyframe *k0 = p->kf_any;
  if (!serial && (p->n_geo || p->n_reproj))
  {
    SAGE_CUDA(cudaEventRecord(ctx->ev_fork, s));
    if (p->n_geo)
      SAGE_CUDA(cudaStreamWaitEvent(sg, ctx->ev_fork, 0));
    if (p->n_reproj)
      SAGE_CUDA(cudaStreamWaitEvent(sr, ctx->ev_fork, 0));
  }
  if (p->n_geo)
  {
    setup_geo_kernel<<<(p->n_geo + 127) / 128, 128, 0, sg>>>(p->geo_d.p, p->geo_ij.p, p->geo_off.p, p->n_geo, poses, codes, scales, p->C,
                                                             p->eps, jac);
    const int HW = p->H * p->W;
    dim3 grid((HW + 255) / 256, p->n_maps);
    {
      ProfScope ps(p, SAGE_BA_PROF_DEPTH_PREP, sg);
      depth_unscaled_batched_kernel<<<grid, 256, 0, sg>>>(p->maps_d.p, codes, HW, p->C);
      depth_pack_batched_kernel<<<grid, 256, 0, sg>>>(p->maps_d.p, p->H, p->W);
    }
    const sage_ba_camera &cam = k0->cams[0];
    ProfScope ps(p, jac ? SAGE_BA_PROF_GEO_JAC : SAGE_BA_PROF_GEO_ERR, sg);
    SAGE_CHECK(launch_geo(jac, p->C, p->geo_d.p, p->n_geo, p->W, p->H, cam.fx, cam.fy, cam.u0, cam.v0, jac ? p->slices_geo : p->slices_geo_err,
                          p->partHg.p, p->partEg.p, out, 1, sg, p->geo_tc) == 0,
               "unsupported code_size");
    ctx->launches += 5;
  }
  if (p->n_reproj)
  {
    setup_reproj_kernel<<<(p->n_reproj + 127) / 128, 128, 0, sr>>>(p->reproj_d.p, p->reproj_ij.p, p->reproj_off.p, p->n_reproj, poses,
                                                                   codes, scales, p->C, p->eps, jac);
    ProfScope ps(p, jac ? SAGE_BA_PROF_REPROJ_JAC : SAGE_BA_PROF_REPROJ_ERR, sr);
    SAGE_CHECK(launch_reproj(jac, false, p->C, p->reproj_d.p, p->n_reproj, out, 1, sr) == 0, "unsupported code_size");
    ctx->launches += 2;
  }
  if (p->n_photo)
  {
    setup_photo_kernel<<<(p->n_photo + 127) / 128, 128, 0, s>>>(p->photo_d.p, p->photo_ij.p, p->photo_off.p, p->n_photo, poses, codes,
                                                                scales, p->C, p->eps, jac);
    ProfScope ps(p, jac ? SAGE_BA_PROF_PHOTO_JAC : SAGE_BA_PROF_PHOTO_ERR);
    SAGE_CHECK(launch_photo(jac ? PH_MAP_JAC : PH_MAP_ERR, p->F, p->C, p->photo_d.p, p->n_photo, k0->pyr, jac ? p->slices_photo : p->slices_photo_err, p->partH.p,
                            p->partE.p, out, 1, 13 + p->C, s, p->staged) == 0,
               "unsupported (feat_channels, code_size)");
    ctx->launches += 3;
  }
  if (!serial)
  {
    if (p->n_geo)
    {
      SAGE_CUDA(cudaEventRecord(ctx->ev_join[0], sg));
      SAGE_CUDA(cudaStreamWaitEvent(s, ctx->ev_join[0], 0));
    }
    if (p->n_reproj)
    {
      SAGE_CUDA(cudaEventRecord(ctx->ev_join[1], sr));
      SAGE_CUDA(cudaStreamWaitEvent(s, ctx->ev_join[1], 0));
    }
  }
  SAGE_CUDA(cudaGetLastError());
}

// make a packed buffer (one segment per owner) complete on every rank
static void exchange(sage_ba_problem *p, float *buf, size_t seg)
{
  if (p->world <= 1)
    return;
  ProfScope ps(p, SAGE_BA_PROF_COMM);
  if (p->comm)
  {
    const std::string e = nccl_allgather_inplace(p->comm->nccl, buf, seg, p->rank, p->ctx->stream);
    SAGE_CHECK(e.empty(), e);
  }
  else if (p->allreduce)
    SAGE_CHECK(p->allreduce(buf, seg * p->world, p->allreduce_user) == 0, "all-reduce callback failed");
}

} // namespace sage

#define SAGE_PTRY(p_)                               \
  sage_ba_problem *prob__ = (p_);                   \
  if (!prob__)                                      \
    return 1;                                       \
  sage_ba_context *ctx__ = prob__->ctx;             \
  try                                               \
  {                                                 \
    SAGE_CUDA(cudaSetDevice(ctx__->device));
#define SAGE_PCATCH                \
  }                                \
  catch (const sage::Error &e)     \
  {                                \
    ctx__->err = e.msg;            \
    return 1;                      \
  }                                \
  catch (const std::exception &e)  \
  {                                \
    ctx__->err = e.what();         \
    return 1;                      \
  }                                \
  return 0;

extern "C" {

int sage_ba_shard_plan(int num_pairs, const int *pair_i, const int *pair_j, int world, int *owner)
{
  if (num_pairs < 0 || (num_pairs > 0 && (!pair_i || !pair_j || !owner)) || world < 1)
    return 1;
  std::vector<std::pair<int, int>> uniq(num_pairs);
  for (int q = 0; q < num_pairs; ++q)
    uniq[q] = {pair_i[q], pair_j[q]};
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  const long P = (long)uniq.size();
  for (int q = 0; q < num_pairs; ++q)
  {
    const long idx = std::lower_bound(uniq.begin(), uniq.end(), std::make_pair(pair_i[q], pair_j[q])) - uniq.begin();
    owner[q] = world <= 1 ? 0 : (int)(idx * world / P);
  }
  return 0;
}

int sage_ba_problem_create(sage_ba_context *ctx, int num_keyframes, sage_ba_keyframe *const *kfs, sage_ba_problem **out)
{
  if (!ctx || !out)
    return 1;
  sage_ba_context *ctx__ = ctx;
  try
  {
    SAGE_CHECK(num_keyframes >= 1 && kfs, "need at least one keyframe");
    sage_ba_problem *p = new sage_ba_problem();
    p->ctx = ctx;
    p->K = num_keyframes;
    for (int k = 0; k < num_keyframes; ++k)
    {
      // a null entry is a keyframe this process never touches (another rank owns every factor it takes part in)
      if (kfs[k])
      {
        if (!p->kf_any)
          p->kf_any = kfs[k];
        SAGE_CHECK(kfs[k]->H == p->kf_any->H && kfs[k]->W == p->kf_any->W && kfs[k]->F == p->kf_any->F && kfs[k]->C == p->kf_any->C &&
                       kfs[k]->L == p->kf_any->L,
                   "keyframes have different shapes");
      }
      p->kfs.push_back(kfs[k]);
    }
    SAGE_CHECK(p->kf_any, "need at least one keyframe");
    if (const char *e = getenv("SAGE_BA_STAGED"))
      p->staged = atoi(e) != 0;
    if (const char *e = getenv("SAGE_BA_DETERMINISTIC"))
      p->deterministic = atoi(e) != 0;
    if (getenv("SAGE_BA_SERIAL_FACTORS"))
      p->concurrent_factors = false;
    p->F = p->kf_any->F;
    p->C = p->kf_any->C;
    p->L = p->kf_any->L;
    p->H = p->kf_any->H;
    p->W = p->kf_any->W;
    p->N = p->kf_any->N;
    SAGE_CUDA(cudaSetDevice(ctx->device));
    for (int w = 0; w < 2; ++w)
    {
      p->state[w][0].ensure((size_t)p->K * 12);
      p->state[w][1].ensure((size_t)p->K * p->C);
      p->state[w][2].ensure((size_t)p->K);
    }
    *out = p;
    return 0;
  }
  catch (const sage::Error &e)
  {
    ctx__->err = e.msg;
    return 1;
  }
}

void sage_ba_problem_destroy(sage_ba_problem *p)
{
  if (!p)
    return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for (void *q : p->owned)
    cudaFree(q);
  for (auto &sp : p->spans)
  {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  for (cudaEvent_t e : p->event_pool)
    cudaEventDestroy(e);
  delete p;
}

static int add_meta(sage_ba_problem *p, int kind, int i, int j, int D)
{
  FactorMeta m;
  m.kind = kind;
  m.i = i;
  m.j = j;
  m.D = D;
  m.off = 0; // offsets are assigned when the problem is built (they depend on the sharding)
  m.cost_off = 0;
  m.owner = -1; // decided when the problem is built (sage_ba_shard_plan over all pairs)
  p->metas.push_back(m);
  p->specs.emplace_back();
  return (int)p->metas.size() - 1;
}

int sage_ba_problem_add_photometric(sage_ba_problem *p, int i, int j, const float *weights)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(i >= 0 && j >= 0 && i < p->K && j < p->K && i != j && weights, "bad keyframe index");
  const int gi = add_meta(p, 0, i, j, 13 + p->C);
  p->specs[gi].w.assign(weights, weights + p->L);
  p->residuals += (long)p->L * p->N * p->F;
  SAGE_PCATCH
}

int sage_ba_problem_add_geometric(sage_ba_problem *p, int i, int j, float loss_param, float weight)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(i >= 0 && j >= 0 && i < p->K && j < p->K && i != j, "bad keyframe index");
  const int gi = add_meta(p, 1, i, j, 14 + 2 * p->C);
  p->specs[gi].loss = loss_param;
  p->specs[gi].weight = weight;
  p->residuals += p->N;
  SAGE_PCATCH
}

int sage_ba_problem_add_reprojection(sage_ba_problem *p, int i, int j, const int32_t *loc1d, const float *homo, const float *match2d,
                                     int M, float loss_param, float weight)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(i >= 0 && j >= 0 && i < p->K && j < p->K && i != j, "bad keyframe index");
  SAGE_CHECK(M > 0 && M <= 4096, "num_matches out of range");
  SAGE_CHECK(loc1d && homo && match2d, "null match arrays");
  for (int m = 0; m < M; ++m)
    SAGE_CHECK(loc1d[m] >= 0 && loc1d[m] < p->H * p->W, "matched location outside the image");
  const int gi = add_meta(p, 2, i, j, 13 + p->C);
  sage_ba_problem::FactorSpec &sp = p->specs[gi];
  sp.loc.assign(loc1d, loc1d + M);
  sp.homo.assign(homo, homo + 3 * (size_t)M);
  sp.uv.assign(match2d, match2d + 2 * (size_t)M);
  sp.loss = loss_param;
  sp.weight = weight;
  p->residuals += 2L * M;
  SAGE_PCATCH
}

int sage_ba_problem_add_code_prior(sage_ba_problem *p, int kf, const float *init_code, float weight)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(kf >= 0 && kf < p->K, "bad keyframe index");
  PriorSpec s;
  memset(&s, 0, sizeof(s));
  s.kind = 0;
  s.kf = kf;
  s.weight = weight;
  if (init_code)
    memcpy(s.init_code, init_code, sizeof(float) * p->C);
  p->priors.push_back(s);
  SAGE_PCATCH
}

int sage_ba_problem_add_scale_prior(sage_ba_problem *p, int kf, float init_scale, float weight)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(kf >= 0 && kf < p->K && init_scale > 0.f, "bad scale prior");
  PriorSpec s;
  memset(&s, 0, sizeof(s));
  s.kind = 1;
  s.kf = kf;
  s.weight = weight;
  s.init_scale = init_scale;
  p->priors.push_back(s);
  SAGE_PCATCH
}

int sage_ba_problem_fix(sage_ba_problem *p, int kf, int fix_pose, int fix_scale)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(kf >= 0 && kf < p->K, "bad keyframe index");
  if (p->fixed_h.empty())
    p->fixed_h.assign(p->dim(), 0);
  if (fix_pose)
    for (int c = 0; c < 6; ++c)
      p->fixed_h[6 * kf + c] = 1;
  if (fix_scale)
    p->fixed_h[6 * p->K + kf * (p->C + 1) + p->C] = 1;
  SAGE_PCATCH
}

int sage_ba_problem_set_solver(sage_ba_problem *p, int solver)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(solver >= 0 && solver <= 2, "solver must be 0 (block Cholesky), 1 (dense cuSOLVER cross-check) or 2 (block Cholesky, natural order)");
  p->solver = solver;
  SAGE_PCATCH
}

int sage_ba_problem_set_shard(sage_ba_problem *p, int rank, int world)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  SAGE_CHECK(world >= 1 && rank >= 0 && rank < world, "bad shard");
  p->rank = rank;
  p->world = world;
  SAGE_PCATCH
}

int sage_ba_problem_set_deterministic(sage_ba_problem *p, int on)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!p->built, "problem already built");
  p->deterministic = on != 0;
  SAGE_PCATCH
}

int sage_ba_problem_set_relinearize_always(sage_ba_problem *p, int always)
{
  if (!p)
    return 1;
  p->relinearize_always = always != 0;
  return 0;
}

int sage_ba_problem_set_state(sage_ba_problem *p, const float *poses, const float *codes, const float *scales, float eps)
{
  SAGE_PTRY(p)
  cudaStream_t s = ctx__->stream;
  p->eps = eps;
  p->lin_valid = false;
  SAGE_CUDA(cudaMemcpyAsync(p->state[0][0].p, poses, sizeof(float) * p->K * 12, cudaMemcpyHostToDevice, s));
  SAGE_CUDA(cudaMemcpyAsync(p->state[0][1].p, codes, sizeof(float) * p->K * p->C, cudaMemcpyHostToDevice, s));
  SAGE_CUDA(cudaMemcpyAsync(p->state[0][2].p, scales, sizeof(float) * p->K, cudaMemcpyHostToDevice, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  SAGE_PCATCH
}

int sage_ba_problem_get_state(sage_ba_problem *p, float *poses, float *codes, float *scales)
{
  SAGE_PTRY(p)
  cudaStream_t s = ctx__->stream;
  SAGE_CUDA(cudaMemcpyAsync(poses, p->state[0][0].p, sizeof(float) * p->K * 12, cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaMemcpyAsync(codes, p->state[0][1].p, sizeof(float) * p->K * p->C, cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaMemcpyAsync(scales, p->state[0][2].p, sizeof(float) * p->K, cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  SAGE_PCATCH
}

int sage_ba_problem_update_map(sage_ba_problem *p, float *poses, float *codes, float *scales, float *dpt_maps, int memory)
{
  SAGE_PTRY(p)
  SAGE_CHECK(memory == SAGE_BA_HOST || memory == SAGE_BA_DEVICE, "memory must be SAGE_BA_HOST or SAGE_BA_DEVICE");
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  const size_t HW = (size_t)p->H * p->W;
  if (poses)
    SAGE_CUDA(cudaMemcpyAsync(poses, p->state[0][0].p, sizeof(float) * p->K * 12, cudaMemcpyDeviceToHost, s));
  if (codes)
    SAGE_CUDA(cudaMemcpyAsync(codes, p->state[0][1].p, sizeof(float) * p->K * p->C, cudaMemcpyDeviceToHost, s));
  if (scales)
    SAGE_CUDA(cudaMemcpyAsync(scales, p->state[0][2].p, sizeof(float) * p->K, cudaMemcpyDeviceToHost, s));
  if (dpt_maps)
  {
    // every keyframe resident on this rank with depth data (all of them unless the caller sharded the uploads)
    int nk = 0;
    for (int k = 0; k < p->K; ++k)
      nk += (p->kfs[k] && p->kfs[k]->bias) ? 1 : 0;
    SAGE_CHECK(nk == p->K, "update_map needs every keyframe's depth data on this rank");
    float *dst = memory == SAGE_BA_DEVICE ? dpt_maps : p->depth_out.ensure((size_t)p->K * HW);
    dim3 grid((unsigned)((HW + 255) / 256), p->K);
    update_depth_batched_kernel<<<grid, 256, 0, s>>>(p->maps_all_d.p, p->state[0][1].p, p->state[0][2].p, (int)HW, p->C, dst);
    ctx__->launches++;
    SAGE_CUDA(cudaGetLastError());
    if (memory == SAGE_BA_HOST)
      SAGE_CUDA(cudaMemcpyAsync(dpt_maps, dst, sizeof(float) * p->K * HW, cudaMemcpyDeviceToHost, s));
  }
  SAGE_CUDA(cudaStreamSynchronize(s));
  SAGE_PCATCH
}

int sage_ba_problem_dim(const sage_ba_problem *p) { return p ? p->dim() : 0; }
int sage_ba_problem_num_factors(const sage_ba_problem *p) { return p ? (int)(p->metas.size() + p->priors.size()) : 0; }
long sage_ba_problem_num_residuals(const sage_ba_problem *p) { return p ? p->residuals : 0; }

int sage_ba_problem_factor_buffer(sage_ba_problem *p, float **ptr, size_t *count)
{
  SAGE_PTRY(p)
  problem_build(p);
  if (ptr)
    *ptr = p->fbuf.p;
  if (count)
    *count = p->fbuf_count;
  SAGE_PCATCH
}

int sage_ba_problem_cost_buffer(sage_ba_problem *p, float **ptr, size_t *count)
{
  SAGE_PTRY(p)
  problem_build(p);
  if (ptr)
    *ptr = p->cbuf.p;
  if (count)
    *count = p->cseg * p->world;
  SAGE_PCATCH
}

int sage_ba_problem_factor_offsets(sage_ba_problem *p, int *offsets, int *cost_offsets, int *owners)
{
  SAGE_PTRY(p)
  problem_build(p);
  for (size_t f = 0; f < p->metas.size(); ++f)
  {
    if (offsets)
      offsets[f] = p->metas[f].off;
    if (cost_offsets)
      cost_offsets[f] = p->metas[f].cost_off;
    if (owners)
      owners[f] = p->metas[f].owner;
  }
  SAGE_PCATCH
}

int sage_ba_problem_solver_info(sage_ba_problem *p, int *num_blocks, int *fill_blocks, int *depth)
{
  SAGE_PTRY(p)
  problem_build(p);
  if (num_blocks)
    *num_blocks = p->bs.nblocks;
  if (fill_blocks)
    *fill_blocks = (int)p->bs.fill_blocks;
  if (depth)
    *depth = p->bs.depth;
  SAGE_PCATCH
}

int sage_ba_problem_solver_trace(sage_ba_problem *p, long long *out, int *positions_to_keyframes)
{
  SAGE_PTRY(p)
  problem_build(p);
  SAGE_CHECK(out, "null output");
  SAGE_CHECK(p->bs.read_trace(out, ctx__->stream) == p->K, "no trace: set SAGE_BA_SOLVER_TRACE=1 before the solve");
  if (positions_to_keyframes)
    for (int q = 0; q < p->K; ++q)
      positions_to_keyframes[q] = p->bs.order[q];
  SAGE_PCATCH
}

int sage_ba_problem_set_allreduce(sage_ba_problem *p, sage_ba_allreduce_fn fn, void *user)
{
  if (!p)
    return 1;
  p->allreduce = fn;
  p->allreduce_user = user;
  return 0;
}

int sage_ba_nccl_unique_id(char id[128])
{
  const std::string e = nccl_unique_id(id);
  if (!e.empty())
    fprintf(stderr, "sage_ba: %s\n", e.c_str());
  return e.empty() ? 0 : 1;
}

int sage_ba_comm_create(sage_ba_context *ctx, const char id[128], int rank, int world, sage_ba_comm **out)
{
  SAGE_TRY(ctx)
  SAGE_CHECK(ctx && out && id && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  SAGE_CUDA(cudaSetDevice(ctx->device));
  void *c = nullptr;
  const std::string e = nccl_comm_create(id, rank, world, &c);
  SAGE_CHECK(e.empty(), e);
  sage_ba_comm *cm = new sage_ba_comm();
  cm->nccl = c;
  cm->rank = rank;
  cm->world = world;
  cm->owned = true;
  *out = cm;
  SAGE_CATCH
}

int sage_ba_comm_wrap(void *nccl_comm, int rank, int world, sage_ba_comm **out)
{
  if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world)
    return 1;
  sage_ba_comm *cm = new sage_ba_comm();
  cm->nccl = nccl_comm;
  cm->rank = rank;
  cm->world = world;
  cm->owned = false;
  *out = cm;
  return 0;
}

void sage_ba_comm_destroy(sage_ba_comm *c)
{
  if (!c)
    return;
  if (c->owned)
    nccl_comm_destroy(c->nccl);
  delete c;
}

int sage_ba_problem_set_comm(sage_ba_problem *p, sage_ba_comm *comm)
{
  SAGE_PTRY(p)
  SAGE_CHECK(!comm || (comm->rank == p->rank && comm->world == p->world), "communicator rank / size differ from the problem's shard");
  p->comm = comm;
  SAGE_PCATCH
}

int sage_ba_problem_exchange(sage_ba_problem *p, int which)
{
  SAGE_PTRY(p)
  problem_build(p);
  if (which == 0)
    exchange(p, p->fbuf.p, p->fseg);
  else
    exchange(p, p->cbuf.p, p->cseg);
  SAGE_PCATCH
}

int sage_ba_problem_linearize(sage_ba_problem *p)
{
  SAGE_PTRY(p)
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  if (p->world > 1 && !p->comm)
    SAGE_CUDA(cudaMemsetAsync(p->fbuf.p, 0, p->fbuf_count * sizeof(float), s)); // sum all-reduce: other ranks' slots must be zero
  run_factors(p, 0, true, p->fbuf.p);
  p->lin_valid = false;
  SAGE_PCATCH
}

int sage_ba_problem_assemble(sage_ba_problem *p, double *H, double *g, double *cost)
{
  SAGE_PTRY(p)
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  const int n = p->dim();
  ProfScope ps(p, SAGE_BA_PROF_ASSEMBLE);
  p->bs.assemble(p->fbuf.p, p->metas_d.p, p->priors_d.p, p->state[0][1].p, p->state[0][2].p, s, &ctx__->launches);
  p->dense_valid = false;
  priors_kernel<<<1, 256, 0, s>>>(p->priors_d.p, (int)p->priors.size(), p->state[0][1].p, p->state[0][2].p, nullptr, nullptr,
                                  p->prior_cost.p, n, p->K, p->C, 0);
  total_cost_kernel<<<1, 256, 0, s>>>(p->fbuf.p, p->errpos_d.p, (int)p->metas.size(), p->prior_cost.p, p->total_cost.p);
  ctx__->launches += 2;
  SAGE_CUDA(cudaGetLastError());
  ps.end();
  p->lin_valid = true;
  if (H || g || cost)
  {
    SAGE_CUDA(cudaMemcpyAsync(p->hcost.p, p->total_cost.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    if (H || g)
      ensure_dense(p);
    if (H)
      SAGE_CUDA(cudaMemcpyAsync(H, p->Hm.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, s));
    if (g)
      SAGE_CUDA(cudaMemcpyAsync(g, p->gv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    SAGE_CUDA(cudaStreamSynchronize(s));
    if (cost)
      *cost = p->hcost.p[0];
  }
  SAGE_PCATCH
}

int sage_ba_problem_solve(sage_ba_problem *p, double damp, double *delta)
{
  SAGE_PTRY(p)
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  const int n = p->dim(), np = 6 * p->K;
  ProfScope ps(p, SAGE_BA_PROF_SOLVE);
  double *dl = p->delta.p;
  SAGE_CUDA(cudaMemsetAsync(p->info.p, 0, sizeof(int) * 4, s));
  if (p->solver != 1)
  {
    // default: block-sparse Cholesky over keyframes (blocksolve.cu) -- per-keyframe Schur complements, one CTA per block column
    p->bs.solve(damp, dl, p->info.p, s, &ctx__->launches);
  }
  else
  {
    // cross-check: the same system densely, Schur complement of the code+scale block fused into ONE cuSOLVER factorisation
    // (that block is ordered first: potrf eliminates it, leaves S = H_pp - H_pc H_cc^-1 H_cp in the trailing 6K x 6K block and
    // factors it; potrs does both substitutions)
    ensure_dense(p);
    p->Hd.ensure((size_t)n * n);
    p->gd.ensure(n);
    if (!p->potrf_lwork)
    {
      SAGE_CHECK(cusolverDnDpotrf_bufferSize(ctx__->cusolver, CUBLAS_FILL_MODE_LOWER, n, p->Hd.p, n, &p->potrf_lwork) == CUSOLVER_STATUS_SUCCESS,
                 "potrf_bufferSize failed");
      p->work.ensure(std::max(p->potrf_lwork, 4));
    }
    damp_perm_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256, 0, s>>>(p->Hm.p, p->gv.p, p->fixed_d.p, p->Hd.p, p->gd.p, n, np, damp);
    SAGE_CHECK(cusolverDnDpotrf(ctx__->cusolver, CUBLAS_FILL_MODE_LOWER, n, p->Hd.p, n, p->work.p, p->potrf_lwork, p->info.p) ==
                   CUSOLVER_STATUS_SUCCESS,
               "potrf(H) failed to launch");
    SAGE_CHECK(cusolverDnDpotrs(ctx__->cusolver, CUBLAS_FILL_MODE_LOWER, n, 1, p->Hd.p, n, p->gd.p, n, p->info.p + 1) == CUSOLVER_STATUS_SUCCESS,
               "potrs failed");
    unpermute_kernel<<<(n + 255) / 256, 256, 0, s>>>(p->gd.p, dl, n, np);
    ctx__->launches += 2;
  }
  retract_kernel<<<(p->K + 63) / 64, 64, 0, s>>>(p->state[0][0].p, p->state[0][1].p, p->state[0][2].p, dl, p->state[1][0].p,
                                                 p->state[1][1].p, p->state[1][2].p, p->K, p->C);
  ctx__->launches += 1;
  SAGE_CUDA(cudaGetLastError());
  ps.end();
  if (delta)
  {
    SAGE_CUDA(cudaMemcpyAsync(delta, dl, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    SAGE_CUDA(cudaMemcpyAsync(p->hinfo.p, p->info.p, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
    SAGE_CUDA(cudaStreamSynchronize(s));
    SAGE_CHECK(p->hinfo.p[0] == 0 && p->hinfo.p[1] == 0, "normal equations are not positive definite");
  }
  SAGE_PCATCH
}

int sage_ba_problem_evaluate(sage_ba_problem *p, int which)
{
  SAGE_PTRY(p)
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  if (p->world > 1 && !p->comm)
    SAGE_CUDA(cudaMemsetAsync(p->cbuf.p, 0, p->cseg * p->world * sizeof(float), s));
  run_factors(p, which ? 1 : 0, false, p->cbuf.p);
  SAGE_PCATCH
}

int sage_ba_problem_cost(sage_ba_problem *p, int which, double *cost)
{
  SAGE_PTRY(p)
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  const int w = which ? 1 : 0;
  priors_kernel<<<1, 256, 0, s>>>(p->priors_d.p, (int)p->priors.size(), p->state[w][1].p, p->state[w][2].p, nullptr, nullptr,
                                  p->prior_cost.p + 1, p->dim(), p->K, p->C, 0);
  total_cost_kernel<<<1, 256, 0, s>>>(p->cbuf.p, p->costpos_d.p, (int)p->metas.size(), p->prior_cost.p + 1, p->total_cost.p + 1);
  ctx__->launches += 2;
  SAGE_CUDA(cudaMemcpyAsync(p->hcost.p + 1, p->total_cost.p + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaMemcpyAsync(p->hinfo.p, p->info.p, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  if (cost)
    *cost = p->hcost.p[1];
  SAGE_PCATCH
}

int sage_ba_problem_accept(sage_ba_problem *p)
{
  SAGE_PTRY(p)
  cudaStream_t s = ctx__->stream;
  for (int q = 0; q < 3; ++q)
  {
    const size_t nfl = q == 0 ? (size_t)p->K * 12 : (q == 1 ? (size_t)p->K * p->C : (size_t)p->K);
    SAGE_CUDA(cudaMemcpyAsync(p->state[0][q].p, p->state[1][q].p, sizeof(float) * nfl, cudaMemcpyDeviceToDevice, s));
  }
  p->lin_valid = false;
  SAGE_PCATCH
}

int sage_ba_problem_profile(sage_ba_problem *p, int enable)
{
  if (!p)
    return 1;
  p->profiling = enable != 0;
  return 0;
}

int sage_ba_problem_profile_read(sage_ba_problem *p, double *ms, long *counts, int reset)
{
  SAGE_PTRY(p)
  SAGE_CUDA(cudaStreamSynchronize(ctx__->stream));
  for (auto &sp : p->spans)
  {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess)
    {
      p->prof_ms[sp.kind] += t;
      p->prof_n[sp.kind] += 1;
    }
    p->event_pool.push_back(sp.a);
    p->event_pool.push_back(sp.b);
  }
  p->spans.clear();
  for (int k = 0; k < SAGE_BA_PROF_KINDS; ++k)
  {
    if (ms)
      ms[k] = p->prof_ms[k];
    if (counts)
      counts[k] = p->prof_n[k];
    if (reset)
    {
      p->prof_ms[k] = 0;
      p->prof_n[k] = 0;
    }
  }
  SAGE_PCATCH
}

int sage_ba_problem_shard_counts(const sage_ba_problem *p, int *n_photo, int *n_geo, int *n_reproj)
{
  if (!p || !p->built)
    return 1;
  if (n_photo)
    *n_photo = p->n_photo;
  if (n_geo)
    *n_geo = p->n_geo;
  if (n_reproj)
    *n_reproj = p->n_reproj;
  return 0;
}

int sage_ba_problem_lm_step(sage_ba_problem *p, double *damp, double min_damp, double max_damp, double damp_dec_factor,
                            double damp_inc_factor, double *cost, double *candidate_cost, int *accepted)
{
  SAGE_PTRY(p)
  SAGE_CHECK(damp, "null damping");
  problem_build(p);
  cudaStream_t s = ctx__->stream;
  // everything of the iteration is enqueued before the host looks at anything: linearise, exchange, assemble, solve, evaluate
  // the candidate, exchange; ONE synchronisation then delivers both costs (the damping of this step does not depend on them).
  // After a rejected step the state has not moved: the linearisation in the factor buffer is still the current one and is
  // reused (the reference's own loop does the same, core/system/camera_tracker.cpp:1159) unless the caller asks for the full
  // iteration every time (bench.py does: BASELINE's metric defines an iteration as including the linearisation).
  if (!p->lin_valid || p->relinearize_always)
  {
    SAGE_CHECK(sage_ba_problem_linearize(p) == 0, ctx__->err);
    exchange(p, p->fbuf.p, p->fseg);
    SAGE_CHECK(sage_ba_problem_assemble(p, nullptr, nullptr, nullptr) == 0, ctx__->err);
  }
  SAGE_CUDA(cudaMemcpyAsync(p->hcost.p, p->total_cost.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  SAGE_CHECK(sage_ba_problem_solve(p, *damp, nullptr) == 0, ctx__->err);
  SAGE_CHECK(sage_ba_problem_evaluate(p, 1) == 0, ctx__->err);
  exchange(p, p->cbuf.p, p->cseg);
  double cand = 0.0;
  SAGE_CHECK(sage_ba_problem_cost(p, 1, &cand) == 0, ctx__->err); // the synchronisation
  const double cur = p->hcost.p[0];
  const bool ok = p->hinfo.p[0] == 0 && p->hinfo.p[1] == 0 && cand < cur;
  if (ok)
  {
    SAGE_CHECK(sage_ba_problem_accept(p) == 0, ctx__->err);
    *damp = std::min(std::max(min_damp, *damp / damp_dec_factor), max_damp);
  }
  else
    *damp = std::min(std::max(min_damp, *damp * damp_inc_factor), max_damp);
  if (cost)
    *cost = cur;
  if (candidate_cost)
    *candidate_cost = cand;
  if (accepted)
    *accepted = ok ? 1 : 0;
  SAGE_PCATCH
}

int sage_ba_problem_lm(sage_ba_problem *p, const sage_ba_lm_options *opt, sage_ba_lm_report *rep)
{
  SAGE_PTRY(p)
  SAGE_CHECK(opt, "null options");
  problem_build(p);
  sage_ba_lm_report r;
  memset(&r, 0, sizeof(r));
  auto clampd = [&](double d) { return std::min(std::max(opt->min_damp, d), opt->max_damp); };
  double damp = opt->init_damp;
  double cost = 0.0;
  SAGE_CHECK(sage_ba_problem_linearize(p) == 0, ctx__->err);
  exchange(p, p->fbuf.p, p->fseg);
  SAGE_CHECK(sage_ba_problem_assemble(p, nullptr, nullptr, &cost) == 0, ctx__->err);
  r.linearizations = 1;
  r.initial_cost = cost;
  for (int it = 0; it < opt->max_iters; ++it)
  {
    r.iterations = it + 1;
    bool improved = false;
    double cand = cost;
    int trials = 0;
    while (true)
    {
      SAGE_CHECK(sage_ba_problem_solve(p, damp, nullptr) == 0, ctx__->err);
      SAGE_CHECK(sage_ba_problem_evaluate(p, 1) == 0, ctx__->err);
      exchange(p, p->cbuf.p, p->cseg);
      SAGE_CHECK(sage_ba_problem_cost(p, 1, &cand) == 0, ctx__->err);
      r.evaluations++;
      const bool solved = p->hinfo.p[0] == 0 && p->hinfo.p[1] == 0;
      if (solved && cand < cost)
      {
        improved = true;
        break;
      }
      if (damp < opt->max_damp && trials < opt->max_trials)
      {
        damp = clampd(damp * opt->damp_inc_factor);
        ++trials;
      }
      else
        break;
    }
    if (!improved)
      break;
    SAGE_CHECK(sage_ba_problem_accept(p) == 0, ctx__->err);
    r.accepted++;
    const double prev = cost;
    cost = cand;
    damp = clampd(damp / opt->damp_dec_factor);
    if (prev - cost < opt->min_rel_decrease * prev || it + 1 >= opt->max_iters)
      break;
    SAGE_CHECK(sage_ba_problem_linearize(p) == 0, ctx__->err);
    exchange(p, p->fbuf.p, p->fseg);
    SAGE_CHECK(sage_ba_problem_assemble(p, nullptr, nullptr, nullptr) == 0, ctx__->err);
    r.linearizations++;
  }
  r.final_cost = cost;
  r.final_damp = damp;
  if (rep)
    *rep = r;
  SAGE_PCATCH
}

} // extern "C"

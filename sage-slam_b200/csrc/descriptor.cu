// descriptor.cu -- dense descriptor cycle-matching, sm_100a.
//
// Replaces the torch expression chain the reference runs in every ReprojectionFactor / MatchGeometryFactor constructor
// and per tracked frame:
//   core/gtsam/reprojection_factor.cpp:57-92, core/gtsam/match_geometry_factor.cpp:62-97,
//   core/system/camera_tracker.cpp:798-834 (FeatureMatchingGeo)
//     response_1[k, p] = -sum_c (desc0[c, kp_k] - desc1[c, p])^2        (materialises [C, K, HW] twice)
//     m_k   = argmax_p response_1[k, p]
//     response_0[k, p] = -sum_c (desc1[c, m_k] - desc0[c, p])^2
//     c_k   = argmax_p response_0[k, p]
//     inlier_k = |pixel(kp_k) - pixel(c_k)|^2 <= thresh^2 ; outputs are the inliers in keypoint order.
//
// Design: thread == keypoint.  A thread keeps its query descriptor in registers and walks the pixels of its CTA's chunk,
// whose descriptors are staged pixel-major in shared memory and read as warp-wide broadcasts (one LDS.128 per 4 channels
// for all 32 keypoints of the warp), so the K x HW response matrix never exists and no cross-thread reduction is needed:
// the running (best, index) lives in two registers.  Arithmetic is the reference's, unfused and in channel order
// (sub, mul, add: no FMA contraction), so the winning INDEX is the one an fp32 evaluation of the expression above gives;
// ties keep the lowest pixel index.  A second tiny kernel merges the per-chunk winners in chunk order and gathers the
// query descriptors of the return pass.  Bound: fp32 issue (3 * CD instructions per keypoint-pixel); bytes are negligible
// (each descriptor map is read once per 256 keypoints).
#include <algorithm>
#include <cstring>

#include "sage_internal.h"

namespace sage
{

constexpr int DM_THREADS = 256; // keypoints per CTA
#ifndef DM_PIX_UNROLL
#define DM_PIX_UNROLL 4 // independent pixels in flight per thread: the channel sum of one pixel is a dependent chain
#endif

// q[k][c] = desc[c][loc[k]]
__global__ void desc_gather_kernel(const float *__restrict__ desc, const int *__restrict__ loc, int K, int CD, int HW, float *__restrict__ q)
{
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * CD)
    return;
  const int k = e / CD, c = e - k * CD;
  q[e] = __ldg(desc + (size_t)c * HW + loc[k]);
}

template <int CD>
__global__ void __launch_bounds__(DM_THREADS)
desc_response_kernel(const float *__restrict__ desc /* [CD][HW] */, const float *__restrict__ q /* [K][CD] */, int K, int HW, int chunk,
                     float *__restrict__ pbest /* [nchunks][K] */, int *__restrict__ pidx)
{
  constexpr int DM_TILE = CD <= 32 ? 256 : 128; // pixels staged per pass (<= 37 KB of shared memory)
  constexpr int kPixUnroll = DM_PIX_UNROLL;
  constexpr int ST = CD + 4; // padded pixel stride: 16-byte aligned rows, 4-way instead of 16-way conflicts on the transposing store
  __shared__ __align__(16) float tile[DM_TILE * ST];
  const int k = blockIdx.y * DM_THREADS + threadIdx.x;
  const int kc = k < K ? k : K - 1;
  float a[CD];
#pragma unroll
  for (int c = 0; c < CD; c += 4)
  {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(q + (size_t)kc * CD + c));
    a[c] = v.x; a[c + 1] = v.y; a[c + 2] = v.z; a[c + 3] = v.w;
  }
  float best = 3.402823466e38f; // smallest sum of squares so far (response = -sum)
  int bidx = 0x7fffffff;
  const int p_begin = blockIdx.x * chunk, p_end = min(p_begin + chunk, HW);
  for (int p0 = p_begin; p0 < p_end; p0 += DM_TILE)
  {
    const int np = min(DM_TILE, p_end - p0);
    __syncthreads();
    for (int e = threadIdx.x; e < CD * DM_TILE; e += DM_THREADS)
    {
      const int c = e / DM_TILE, p = e - c * DM_TILE;
      if (p < np)
        tile[p * ST + c] = __ldg(desc + (size_t)c * HW + p0 + p);
    }
    __syncthreads();
#pragma unroll kPixUnroll
    for (int p = 0; p < np; ++p)
    {
      const float *b = tile + p * ST;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < CD; c += 4)
      {
        const float4 v = *reinterpret_cast<const float4 *>(b + c);
        const float d0 = __fsub_rn(a[c], v.x), d1 = __fsub_rn(a[c + 1], v.y), d2 = __fsub_rn(a[c + 2], v.z), d3 = __fsub_rn(a[c + 3], v.w);
        acc = __fadd_rn(acc, __fmul_rn(d0, d0));
        acc = __fadd_rn(acc, __fmul_rn(d1, d1));
        acc = __fadd_rn(acc, __fmul_rn(d2, d2));
        acc = __fadd_rn(acc, __fmul_rn(d3, d3));
      }
      if (acc < best) // strict: the first (lowest) pixel index wins ties
      {
        best = acc;
        bidx = p0 + p;
      }
    }
  }
  if (k < K)
  {
    pbest[(size_t)blockIdx.x * K + k] = best;
    pidx[(size_t)blockIdx.x * K + k] = bidx;
  }
}

// merge the per-chunk winners in chunk (= pixel) order; optionally gather the winner's descriptor from `desc` as the next query
__global__ void desc_merge_kernel(const float *__restrict__ pbest, const int *__restrict__ pidx, int nchunks, int K, int *__restrict__ arg,
                                  const float *__restrict__ desc, int CD, int HW, float *__restrict__ qnext)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K)
    return;
  float best = pbest[k];
  int bidx = pidx[k];
  for (int s = 1; s < nchunks; ++s)
  {
    const float v = pbest[(size_t)s * K + k];
    if (v < best)
    {
      best = v;
      bidx = pidx[(size_t)s * K + k];
    }
  }
  arg[k] = bidx;
  if (qnext)
    for (int c = 0; c < CD; ++c)
      qnext[(size_t)k * CD + c] = __ldg(desc + (size_t)c * HW + bidx);
}

// cycle-consistency test and order-preserving compaction (torch::nonzero), one CTA.
//   out_sel[m]  = position (0..K-1) of the m-th inlier keypoint ; out_count = M
__global__ void desc_cycle_select_kernel(const int *__restrict__ kp_loc, const int *__restrict__ cyc_loc, int K, int W, float thresh_sq,
                                         int *__restrict__ out_sel, int *__restrict__ out_count)
{
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0)
    base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k0 = 0; k0 < K; k0 += blockDim.x)
  {
    const int k = k0 + threadIdx.x;
    bool in = false;
    if (k < K)
    {
      const int a = kp_loc[k], b = cyc_loc[k];
      // fmod(loc, W) / floor(loc / W) of the reference are exact for loc < 2^24; same integers here
      const float dx = (float)(a % W - b % W), dy = (float)(a / W - b / W);
      in = dx * dx + dy * dy <= thresh_sq;
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (lane == 0)
      warp_tot[wid] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < wid; ++w)
      off += warp_tot[w];
    if (in)
      out_sel[off + __popc(m & ((1u << lane) - 1u))] = k;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      int t = 0;
      for (int w = 0; w < nw; ++w)
        t += warp_tot[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *out_count = base;
}

template <int CD>
static void launch_response(const float *desc, const float *q, int K, int HW, int chunk, int nchunks, float *pbest, int *pidx, cudaStream_t s)
{
  dim3 grid(nchunks, (K + DM_THREADS - 1) / DM_THREADS);
  desc_response_kernel<CD><<<grid, DM_THREADS, 0, s>>>(desc, q, K, HW, chunk, pbest, pidx);
}

static void launch_response_any(int CD, const float *desc, const float *q, int K, int HW, int chunk, int nchunks, float *pbest, int *pidx,
                                cudaStream_t s)
{
  switch (CD)
  {
  case 8: launch_response<8>(desc, q, K, HW, chunk, nchunks, pbest, pidx, s); break;
  case 16: launch_response<16>(desc, q, K, HW, chunk, nchunks, pbest, pidx, s); break;
  case 32: launch_response<32>(desc, q, K, HW, chunk, nchunks, pbest, pidx, s); break;
  case 64: launch_response<64>(desc, q, K, HW, chunk, nchunks, pbest, pidx, s); break;
  default: throw Error{"descriptor channels must be 8, 16, 32 or 64"};
  }
}

} // namespace sage



using namespace sage;

extern "C" int sage_ba_cycle_match(sage_ba_context *ctx, int memory, const float *feat_desc_0, const float *feat_desc_1, int channels, int height,
                                   int width, const int64_t *keypoint_locations_1d, int num_keypoints, float cyc_consis_thresh,
                                   int32_t *raw_matched_locations_1d_1, int32_t *cyc_matched_locations_1d_0, int32_t *inlier_positions,
                                   int *num_inliers, float *kernel_ms)
{
  SAGE_TRY(ctx)
  SAGE_CHECK(ctx, "null context");
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(feat_desc_0 && feat_desc_1 && keypoint_locations_1d && num_inliers, "null argument");
  SAGE_CHECK(memory == SAGE_BA_HOST || memory == SAGE_BA_DEVICE, "memory must be SAGE_BA_HOST or SAGE_BA_DEVICE");
  const int K = num_keypoints, CD = channels, HW = height * width;
  SAGE_CHECK(CD == 8 || CD == 16 || CD == 32 || CD == 64, "descriptor channels must be 8, 16, 32 or 64");
  SAGE_CHECK(K > 0 && K <= (1 << 20), "num_keypoints out of range");
  SAGE_CHECK(height > 0 && width > 0 && (long)height * width < (1l << 24), "image size out of range");
  for (int k = 0; k < K; ++k)
    SAGE_CHECK(keypoint_locations_1d[k] >= 0 && keypoint_locations_1d[k] < HW, "keypoint location out of range");
  cudaStream_t s = ctx->stream;

  // chunking: ~2 CTAs per SM per 256 keypoints, a whole number of staging tiles per chunk
  const int ky = (K + DM_THREADS - 1) / DM_THREADS;
  int nchunks = std::max(1, (2 * ctx->num_sms + ky - 1) / ky);
  int chunk = (HW + nchunks - 1) / nchunks;
  chunk = ((chunk + 31) / 32) * 32;
  nchunks = (HW + chunk - 1) / chunk;

  const size_t map_floats = (size_t)CD * HW;
  const float *d0 = feat_desc_0, *d1 = feat_desc_1;
  if (memory == SAGE_BA_HOST)
  {
    float *dm = ctx->dm_maps.ensure(2 * map_floats);
    SAGE_CUDA(cudaMemcpyAsync(dm, feat_desc_0, map_floats * sizeof(float), cudaMemcpyHostToDevice, s));
    SAGE_CUDA(cudaMemcpyAsync(dm + map_floats, feat_desc_1, map_floats * sizeof(float), cudaMemcpyHostToDevice, s));
    d0 = dm;
    d1 = dm + map_floats;
  }
  int *ibuf = ctx->dm_int.ensure((size_t)K * 4 + 1 + (size_t)nchunks * K);
  int *kp = ibuf, *raw = ibuf + K, *cyc = ibuf + 2 * K, *sel = ibuf + 3 * K, *cnt = ibuf + 4 * K, *pidx = ibuf + 4 * K + 1;
  float *fbuf = ctx->dm_float.ensure((size_t)2 * K * CD + (size_t)nchunks * K);
  float *q0 = fbuf, *q1 = fbuf + (size_t)K * CD, *pbest = fbuf + (size_t)2 * K * CD;
  int *hk = ctx->dm_hint.ensure((size_t)K * 3 + 1);
  for (int k = 0; k < K; ++k)
    hk[k] = (int)keypoint_locations_1d[k];
  SAGE_CUDA(cudaMemcpyAsync(kp, hk, sizeof(int) * K, cudaMemcpyHostToDevice, s));

  struct Events // destroyed on every exit path
  {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events()
    {
      if (a)
        cudaEventDestroy(a);
      if (b)
        cudaEventDestroy(b);
    }
  } ev;
  if (kernel_ms)
  {
    SAGE_CUDA(cudaEventCreate(&ev.a));
    SAGE_CUDA(cudaEventCreate(&ev.b));
    SAGE_CUDA(cudaEventRecord(ev.a, s));
  }
  desc_gather_kernel<<<(K * CD + 255) / 256, 256, 0, s>>>(d0, kp, K, CD, HW, q0);
  launch_response_any(CD, d1, q0, K, HW, chunk, nchunks, pbest, pidx, s);
  desc_merge_kernel<<<(K + 127) / 128, 128, 0, s>>>(pbest, pidx, nchunks, K, raw, d1, CD, HW, q1);
  launch_response_any(CD, d0, q1, K, HW, chunk, nchunks, pbest, pidx, s);
  desc_merge_kernel<<<(K + 127) / 128, 128, 0, s>>>(pbest, pidx, nchunks, K, cyc, nullptr, CD, HW, nullptr);
  desc_cycle_select_kernel<<<1, 1024, 0, s>>>(kp, cyc, K, width, cyc_consis_thresh * cyc_consis_thresh, sel, cnt);
  ctx->launches += 6;
  if (kernel_ms)
    SAGE_CUDA(cudaEventRecord(ev.b, s));
  SAGE_CUDA(cudaGetLastError());
  // raw | cyc | sel | count are contiguous
  SAGE_CUDA(cudaMemcpyAsync(hk, raw, sizeof(int) * ((size_t)3 * K + 1), cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  if (kernel_ms)
    SAGE_CUDA(cudaEventElapsedTime(kernel_ms, ev.a, ev.b));
  const int M = hk[3 * K];
  if (raw_matched_locations_1d_1)
    memcpy(raw_matched_locations_1d_1, hk, sizeof(int) * K);
  if (cyc_matched_locations_1d_0)
    memcpy(cyc_matched_locations_1d_0, hk + K, sizeof(int) * K);
  if (inlier_positions)
    memcpy(inlier_positions, hk + 2 * K, sizeof(int) * M);
  *num_inliers = M;
  SAGE_CATCH
}

// photometric.cu -- fused photometric linearisation / error kernels (sm_100a).
//
// Replaces, for every pyramid level at once and without materialising the Jacobian:
//   photometric_jac_error_calculate_kernel   cuda/photometric_factor_kernels.cpp:33-368   (PH_MAP_JAC)
//   photometric_error_calculate_kernel       :370-522                                    (PH_MAP_ERR)
//   tracker_photo_jac_error_calculate_kernel :524-695  / _with_scale_kernel :697-873     (PH_TRK_JAC)
//   tracker_photo_error_calculate_kernel     :875-988                                    (PH_TRK_ERR)
// plus the ATen reductions that follow them (:1139-1161, :1220-1242, :1301-1322, :1049-1057).
//
// Design (one warp = one independent worker, no block barriers in the main loop):
//  * A warp takes batches of 32 sample points.  Warp geometry, mask lookup and the bilinear tap sets of
//    every level are computed once per sample with lane == sample.
//  * The gathers run with lane == channel quad: a group of F/4 lanes fetches one sample's taps as float4
//    from the channel-last pyramid [SP][3][F] (feature | d/dx | d/dy), so every tap of every map is one
//    fully used 128-byte line for F = 32.  The owner lane broadcasts the packed tap descriptor and the
//    four (zero-padded) weights with 5 shuffles per map set; out-of-bounds taps have weight 0 and a
//    clamped address, so the loads need no predicates.
//  * The F x L residual rows of a sample collapse to the 2x2 Gram matrix G = sum_l w_l sum_ch g~ g~^T and
//    b = sum g~ r (g~ = level-scaled sampled gradient): J_row = g~^T P^ with P^ (2 x D) independent of
//    channel and level.  The six partial sums are reduce-scattered inside the group (7 shuffles) and
//    accumulated over levels; a Cholesky factor of G turns them into two "virtual rows" P^T L of width
//    8 + C ([pose0 6 | scale | rhs | code C]; the pose1 block is exactly -pose0 and is expanded at the end).
//  * J^T J | J^T r of the staged rows is accumulated per warp on the tensor cores: mma.sync m16n8k8 TF32 with
//    the 3xTF32 split (hi*hi + hi*lo + lo*hi, fp32 accumulate), upper-triangular tiles only, operand
//    fragments shared between the A and B roles because both are the same staged rows.
//  * Each CTA writes one private partial; a tiny second kernel reduces partials in a fixed order
//    (deterministic) and applies the inlier normalisation / zero-overlap fallback of the reference.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

#ifndef PH_NWARPS
#define PH_NWARPS 4
#endif
#ifndef PH_HALF
#define PH_HALF 0
#endif
constexpr int PH_WARPS = PH_NWARPS;
constexpr int PH_CTA = PH_WARPS * 32;

template <int F, int C, int MODE>
struct PhotoTraits
{
  static constexpr bool kJac = (MODE == PH_MAP_JAC || MODE == PH_TRK_JAC);
  static constexpr bool kMap = (MODE == PH_MAP_JAC || MODE == PH_MAP_ERR);
  static constexpr int LPG = F / 4;     // lanes per sample group (8 for F = 32, 4 for F = 16)
  static constexpr int NG = 32 / LPG;   // groups per warp
  static constexpr int VPL = 8 / LPG;   // reduced values each lane ends up owning
  static constexpr int WP = kMap ? 8 + C : 8; // staged row: [pose 6 | scale | rhs | code C]
  static constexpr int ST = MmaSyrk<WP>::ST;
  static constexpr int CCH = (C / 4 + LPG - 1) / LPG; // code chunks (float4) per lane
};

// reduce-scatter of 8 values over a group of LPG lanes: lane gl ends with VPL = 8/LPG totals, value index gl*VPL+k
template <int LPG>
__device__ __forceinline__ void reduce_scatter8(float (&v)[8], int gl, float (&out)[8 / LPG])
{
  static_assert(LPG == 8 || LPG == 4, "LPG must be 4 or 8");
  if constexpr (LPG == 8)
  {
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      const bool hi = gl & 4;
      const float send = hi ? v[k] : v[k + 4];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
      a[k] = (hi ? v[k + 4] : v[k]) + recv;
    }
    float b[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
      const bool hi = gl & 2;
      const float send = hi ? a[k] : a[k + 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
      b[k] = (hi ? a[k + 2] : a[k]) + recv;
    }
    const bool hi = gl & 1;
    const float send = hi ? b[0] : b[1];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    out[0] = (hi ? b[1] : b[0]) + recv;
  }
  else
  {
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      const bool hi = gl & 2;
      const float send = hi ? v[k] : v[k + 4];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
      a[k] = (hi ? v[k + 4] : v[k]) + recv;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
      const bool hi = gl & 1;
      const float send = hi ? a[k] : a[k + 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
      out[k] = (hi ? a[k + 2] : a[k]) + recv;
    }
  }
}

#ifndef PH_MINB
#define PH_MINB 3
#endif
#ifndef PH_MINB_ERR
#define PH_MINB_ERR 4 // error-only kernels: 128 registers, 16 warps per SM
#endif
#ifndef PH_UNROLL
#define PH_UNROLL 4
#endif
#ifndef PH_UNROLL_ERR
#define PH_UNROLL_ERR 8 // error-only kernels fetch 5 lines per sample-level instead of 13: twice the samples in flight (2.67 -> 2.50 ms)
#endif
#ifndef PH_FLUSH
#define PH_FLUSH 4 // batches of 32 samples per warp between two flushes of the tensor-core accumulators (MmaSyrk::flush_tiles)
#endif
#ifndef PH_ILV
#define PH_ILV 1 // gather step i serves samples i*NG + q (x-adjacent samples in ONE instruction: their taps share 128-byte lines at levels >= 1)
#endif
#define PH_STR2(x) #x
#define PH_STR(x) PH_STR2(x)

// ------------------------------------------------------------------------------------------------
// Staged variant (STG): the taps of pyramid levels >= 1 come from shared memory.
// A CTA's PH_WARPS warps work on one TILE of PH_WARPS consecutive batches at a time; with the tile-major sample order the
// library builds for problem keyframes (api.cu: sort_samples) that is a 32 x PH_WARPS block of pixels, whose projection into
// frame 1 at level l covers about (32/2^l + 2) x (PH_WARPS/2^l + 2) pixels.  Each warp reduces the bounding box of its taps
// per level (REDUX), one block barrier publishes the boxes, and warp 0 copies the window rows of fg1 -- contiguous runs of
// [3][F] floats in the channel-last layout -- into shared memory with cp.async.bulk (TMA engine, SASS UBLKCP), completion
// counted on an mbarrier (expect_tx).  The copy flies while every warp gathers level 0 straight from global memory (no reuse
// there beyond what L1 gives); the levels above, where 4 / 16 / 64 samples share a cell, then read their 12 lines per
// sample-level from the window at shared-memory latency instead of L2 latency, with no register held per line in flight.
// Taps outside the window (clipped box: strong rotation, depth discontinuity) keep their global address: generic loads
// serve both, there is no second code path.
// ------------------------------------------------------------------------------------------------
#ifndef PH_WIN_KB_JAC
#define PH_WIN_KB_JAC 48 // window bytes per CTA: with the row staging and the second-level accumulators two CTAs fit an SM
#endif
#ifndef PH_WIN_KB_ERR
#define PH_WIN_KB_ERR 52 // error-only kernels have no row staging: 4 CTAs/SM
#endif
#ifndef PH_MINB_STG
#define PH_MINB_STG (PH_HALF ? 3 : 2)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct WinBox
{
  int x0, y0, w, h, base; // window origin / size in pixels of the level, first float of the window in the staging area
};

// window of level l from the PH_WARPS per-warp boxes (xmin, ymin, xmax, ymax), clipped to what is left of the budget
__device__ __forceinline__ WinBox make_window(const int (*box)[4], int nwarps, int used_px, int budget_px, int stride)
{
  int xmin = 0x7fffffff, ymin = 0x7fffffff, xmax = -1, ymax = -1;
  for (int w = 0; w < nwarps; ++w)
  {
    xmin = min(xmin, box[w][0]);
    ymin = min(ymin, box[w][1]);
    xmax = max(xmax, box[w][2]);
    ymax = max(ymax, box[w][3]);
  }
  WinBox b;
  b.x0 = xmin;
  b.y0 = ymin;
  b.w = xmax >= xmin ? xmax - xmin + 1 : 0;
  b.h = ymax >= ymin ? ymax - ymin + 1 : 0;
  const int avail = budget_px - used_px;
  if (b.w > avail)
    b.w = avail;
  if (b.w > 0 && b.w * b.h > avail)
    b.h = avail / b.w;
  if (b.h > 32)
    b.h = 32; // one lane of warp 0 issues one row
  if (b.w <= 0 || b.h <= 0)
    b.w = b.h = 0;
  b.base = used_px * stride;
  return b;
}

template <int F, int C, int MODE, bool STG>
__global__ void __launch_bounds__(PH_CTA, STG ? ((MODE == PH_MAP_JAC || MODE == PH_TRK_JAC) ? PH_MINB_STG : PH_MINB_ERR)
                                              : ((MODE == PH_MAP_JAC || MODE == PH_TRK_JAC) ? PH_MINB : PH_MINB_ERR))
photo_kernel(const PhotoFactor *__restrict__ factors, const __grid_constant__ CamPyr cam, float *__restrict__ partH,
             float *__restrict__ partE)
{
  using T = PhotoTraits<F, C, MODE>;
  constexpr int LPG = T::LPG, NG = T::NG, VPL = T::VPL, WP = T::WP, ST = T::ST, CCH = T::CCH;
  constexpr int kUnroll = T::kJac ? PH_UNROLL : PH_UNROLL_ERR; // gather iterations whose loads are issued together
  constexpr int ROWS = PH_HALF ? 32 : 64; // virtual rows staged per warp at a time (two per sample, 32 samples per batch)
  constexpr int STAGE = T::kJac ? PH_WARPS * ROWS * ST : 4;
  constexpr int HS = T::kJac ? WP * WP : 4;
  __shared__ __align__(16) float Y[STAGE > HS ? STAGE : HS];
  __shared__ __align__(16) float XP[PH_WARPS][32 * 9 + 3];
  __shared__ PhotoFactor fs;
  __shared__ float red[32];
  // staged variant: window (dynamic shared memory), its mbarrier and the per-warp tap boxes of levels 1..L-1
  extern __shared__ __align__(128) float win[];
  __shared__ __align__(8) uint64_t win_bar;
  __shared__ int wbox[STG ? SAGE_MAX_LEVELS : 1][PH_WARPS][4];
  constexpr int WIN_PX = STG ? ((T::kJac ? PH_WIN_KB_JAC : PH_WIN_KB_ERR) * 1024) / (3 * F * 4) : 0;
  uint32_t win_phase = 0;
  if constexpr (STG)
  {
    if (threadIdx.x == 0)
    {
      mbar_init(&win_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }

  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.y);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(PhotoFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPG, q = lane / LPG; // lane inside its group, group inside the warp
  const int N = fs.N, L = cam.L;
  float *Yw = Y + (size_t)warp * ROWS * ST;
  float *xp = XP[warp];

  MmaSyrk<WP> syrk;
  if constexpr (T::kJac)
    syrk.init();
  float err_acc = 0.f, inl_acc = 0.f;
  // second-level accumulator of this warp (fragment order, dynamic shared memory behind the staged variant's window): the
  // tensor-core accumulators are folded into it every PH_FLUSH batches, see MmaSyrk::flush_frag
  constexpr int kAccF4 = MmaSyrk<WP>::NTILES * 32;
  float4 *acc_s = reinterpret_cast<float4 *>(win + (STG ? WIN_PX * 3 * F : 0));
  float4 *acc_w = acc_s + (size_t)warp * kAccF4;
  int nacc = 0;
  if constexpr (T::kJac)
  {
    for (int i = lane; i < kAccF4; i += 32)
      acc_w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }
  (void)acc_w;

  // per-level scale of the value(s) this lane accumulates: [Gxx Gxy Gyy bx by e - -] -> w_l * {fx fx, fx fy, fy fy, fx, fy, 1}
  int selA[VPL], selB[VPL]; // 0: 1, 1: fx_l, 2: fy_l
#pragma unroll
  for (int k = 0; k < VPL; ++k)
  {
    const int vi = gl * VPL + k;
    selA[k] = (vi == 0 || vi == 1 || vi == 3) ? 1 : ((vi == 2 || vi == 4) ? 2 : 0);
    selB[k] = vi == 0 ? 1 : ((vi == 1 || vi == 2) ? 2 : 0);
  }

  const int nbatch = (N + 31) / 32;
  // STG: every warp of the CTA runs the same number of rounds (block barrier inside); a warp past the end works on masked lanes
  const int nround = STG ? (nbatch + PH_WARPS - 1) / PH_WARPS : nbatch;
  for (int it = STG ? blockIdx.x : blockIdx.x * PH_WARPS + warp; it < nround; it += STG ? gridDim.x : gridDim.x * PH_WARPS)
  {
    const int batch = STG ? it * PH_WARPS + warp : it;
    // ------------------------------------------------------------------ lane == sample: geometry
    const int n = batch * 32 + lane;
    const bool live = n < N;
    const int nc = live ? n : N - 1; // clamped index: dead lanes read valid memory and are masked out below
    const float4 hm = __ldg(fs.homo + nc);
    const float hx = hm.x, hy = hm.y, hz = hm.z;
    int idx = 0;
    float d0;
    if constexpr (T::kMap)
    {
      idx = __ldg(fs.loc1d + nc);
      // sampled_dpts_0 = scale_0 * (bias[idx] + jac[idx,:] . code)   (:1094-1095); the dot runs lane == channel quad
      float mydot = 0.f;
#pragma unroll
      for (int i = 0; i < LPG; ++i)
      {
        const int sidx = __shfl_sync(0xffffffffu, idx, q * LPG + i);
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < CCH; ++k)
        {
          const int ch = gl + k * LPG;
          if (ch < C / 4)
          {
            const float4 cb = ldg4(fs.basis0 + (size_t)sidx * C + ch * 4);
            dot += cb.x * fs.code0[ch * 4 + 0] + cb.y * fs.code0[ch * 4 + 1] + cb.z * fs.code0[ch * 4 + 2] + cb.w * fs.code0[ch * 4 + 3];
          }
        }
        dot = group_sum<LPG>(dot);
        if (gl == i)
          mydot = dot;
      }
      d0 = fs.scale0 * (__ldg(fs.bias0 + idx) + mydot);
    }
    else
      d0 = fs.dmul * __ldg(fs.dpts0 + nc);
    const float rx = fs.R10[0] * hx + fs.R10[1] * hy + fs.R10[2] * hz;
    const float ry = fs.R10[3] * hx + fs.R10[4] * hy + fs.R10[5] * hz;
    const float rz = fs.R10[6] * hx + fs.R10[7] * hy + fs.R10[8] * hz;
    const float px0 = d0 * rx + fs.t10[0], py0 = d0 * ry + fs.t10[1], pz0 = d0 * rz + fs.t10[2];
    const bool pos = pz0 > fs.eps;
    float ux = (px0 / pz0) * cam.ofx + cam.ocx;
    float uy = (py0 / pz0) * cam.ofy + cam.ocy;
    // nearest lookup in the full-resolution mask (:158-166); CUDA round() == roundf()
    const int mx = (int)roundf(ux), my = (int)roundf(uy);
    const float wm = (live && pos && within(mx, my, cam.ow, cam.oh)) ? __ldg(fs.mask1 + my * cam.ow + mx) : 0.f;
    const float valid = wm;
    if (valid == 0.f)
    {
      ux = 0.f; // keep every later quantity finite: the sample is multiplied by valid == 0 at the end
      uy = 0.f;
    }
    // JAC: value (gl*VPL+k) of sample (q*LPG+i) is accumulated over levels in xp[sample][value] by this lane only
    float eacc = 0.f; // ERR: error of my own sample
    if constexpr (T::kJac)
    {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        xp[lane * 9 + k] = 0.f;
      __syncwarp();
    }

    if constexpr (STG)
    {
      // bounding box of this warp's clamped taps at every level >= 1 (valid samples only)
      for (int l = 1; l < L; ++l)
      {
        const int W = cam.w[l], H = cam.h[l];
        const float pxl = (ux + 0.5f) * cam.fx[l] / cam.ofx - 0.5f, pyl = (uy + 0.5f) * cam.fy[l] / cam.ofy - 0.5f;
        const int x0 = (int)floorf(pxl), y0 = (int)floorf(pyl);
        const int x1 = x0 < 0x7fffffff ? x0 + 1 : x0, y1 = y0 < 0x7fffffff ? y0 + 1 : y0;
        const int xa = min(max(x0, 0), W - 1), xb = min(max(x1, 0), W - 1), ya = min(max(y0, 0), H - 1), yb = min(max(y1, 0), H - 1);
        const bool in = valid != 0.f;
        const int bx0 = __reduce_min_sync(0xffffffffu, in ? xa : 0x7fffffff), by0 = __reduce_min_sync(0xffffffffu, in ? ya : 0x7fffffff);
        const int bx1 = __reduce_max_sync(0xffffffffu, in ? xb : -1), by1 = __reduce_max_sync(0xffffffffu, in ? yb : -1);
        if (lane == 0)
        {
          wbox[l][warp][0] = bx0;
          wbox[l][warp][1] = by0;
          wbox[l][warp][2] = bx1;
          wbox[l][warp][3] = by1;
        }
      }
      __syncthreads(); // boxes visible; every warp is done reading the previous tile's window
      if (warp == 0)
      {
        int used = 0;
        uint32_t bytes = 0;
        for (int l = 1; l < L; ++l)
        {
          const WinBox b = make_window(wbox[l], PH_WARPS, used, WIN_PX, 3 * F);
          used += b.w * b.h;
          bytes += (uint32_t)(b.w * b.h) * (3 * F * 4);
        }
        if (lane == 0)
        {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic reads of the old window before the async writes
          mbar_expect_tx(&win_bar, bytes);
        }
        __syncwarp();
        used = 0;
        for (int l = 1; l < L; ++l)
        {
          const WinBox b = make_window(wbox[l], PH_WARPS, used, WIN_PX, 3 * F);
          used += b.w * b.h;
          if (lane < b.h)
            bulk_g2s(win + b.base + (size_t)lane * b.w * (3 * F),
                     fs.fg1 + ((size_t)cam.off[l] + (size_t)(b.y0 + lane) * cam.w[l] + b.x0) * (3 * F), (uint32_t)b.w * (3 * F * 4), &win_bar);
        }
      }
    }
    int win_used = 0;

    for (int l = 0; l < L; ++l)
    {
      const int W = cam.w[l], H = cam.h[l];
      // pixel at level l = (pixel_0 + 0.5) * f_l / f_0 - 0.5   (:142-144)
      TapSet t1 = make_tapset((ux + 0.5f) * cam.fx[l] / cam.ofx - 0.5f, (uy + 0.5f) * cam.fy[l] / cam.ofy - 0.5f, W, H, 3 * F);
      int win_rowo = 0;
      const float *wbase = nullptr;
      if constexpr (STG)
      {
        if (l >= 1)
        {
          if (l == 1)
          {
            mbar_wait(&win_bar, win_phase);
            win_phase ^= 1u;
          }
          const WinBox b = make_window(wbox[l], PH_WARPS, win_used, WIN_PX, 3 * F);
          win_used += b.w * b.h;
          win_rowo = b.w * (3 * F);
          wbase = win + b.base + gl * 4;
          // re-express the tap set relative to the window when all four (clamped) taps lie inside it: bit 2 of pk marks it
          const int pix = t1.pk / (3 * F); // clamped nw pixel (the flags live below 3F)
          const int ya = pix / W, xa = pix - ya * W;
          const int xb = xa + ((t1.pk & 2) ? 1 : 0), yb = ya + ((t1.pk & 1) ? 1 : 0);
          if (xa >= b.x0 && xb < b.x0 + b.w && ya >= b.y0 && yb < b.y0 + b.h)
            t1.pk = (((ya - b.y0) * b.w + (xa - b.x0)) * (3 * F)) | 4 | (t1.pk & 3);
        }
      }
      const float *fg1 = fs.fg1 + (size_t)cam.off[l] * (3 * F) + gl * 4;
      const float *sf0 = fs.sfeat0 + (size_t)l * N * F + gl * 4; // KF0 features pre-sampled at its own sample points
      const int rowo = W * (3 * F);
      float lsc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        lsc[k] = fs.w[l] * (selA[k] == 1 ? cam.fx[l] : selA[k] == 2 ? cam.fy[l] : 1.f) *
                 (selB[k] == 1 ? cam.fx[l] : selB[k] == 2 ? cam.fy[l] : 1.f);

      // -------------------------------------------------------------- lane == channel quad: gathers
      float ev[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; // ERR: squared error of sample i of my group, my channel quad
#pragma unroll kUnroll
      for (int i = 0; i < LPG; ++i)
      {
        // which sample the group serves in step i.  Interleaved (PH_ILV): the NG groups of the warp take NG raster-adjacent
        // samples, so at pyramid level l >= 1 (2^l samples per cell and row) one LDG touches 1-3 distinct lines instead of NG.
        const int src = PH_ILV ? i * NG + q : q * LPG + i;
        const TapSet s1 = shfl_tapset(t1, src);
        const bool inw = STG && (s1.pk & 4);
        const float *pnw = (inw ? wbase : fg1) + (s1.pk & ~7);
        const float *pne = pnw + ((s1.pk & 2) ? 3 * F : 0);
        const float *psw = pnw + ((s1.pk & 1) ? (inw ? win_rowo : rowo) : 0);
        const float *pse = psw + ((s1.pk & 2) ? 3 * F : 0);
        const float4 f1 = gather4<STG>(pnw, pse, psw, pne, s1.w);
        const int ns = min(batch * 32 + src, N - 1);
        const float4 f0 = ldg4(sf0 + (size_t)ns * F);
        const float dfx = f0.x - f1.x, dfy = f0.y - f1.y, dfz = f0.z - f1.z, dfw = f0.w - f1.w;
        const float e = dfx * dfx + dfy * dfy + dfz * dfz + dfw * dfw;
        if constexpr (T::kJac)
        {
          const float4 gx = gather4<STG>(pnw + F, pse + F, psw + F, pne + F, s1.w);
          const float4 gy = gather4<STG>(pnw + 2 * F, pse + 2 * F, psw + 2 * F, pne + 2 * F, s1.w);
          float v[8];
          v[0] = gx.x * gx.x + gx.y * gx.y + gx.z * gx.z + gx.w * gx.w;
          v[1] = gx.x * gy.x + gx.y * gy.y + gx.z * gy.z + gx.w * gy.w;
          v[2] = gy.x * gy.x + gy.y * gy.y + gy.z * gy.z + gy.w * gy.w;
          v[3] = gx.x * dfx + gx.y * dfy + gx.z * dfz + gx.w * dfw;
          v[4] = gy.x * dfx + gy.y * dfy + gy.z * dfz + gy.w * dfw;
          v[5] = e;
          v[6] = 0.f;
          v[7] = 0.f;
          float r[VPL];
          reduce_scatter8<LPG>(v, gl, r);
#pragma unroll
          for (int k = 0; k < VPL; ++k)
          {
            float *dstv = xp + src * 9 + gl * VPL + k;
            *dstv = fmaf(lsc[k], r[k], *dstv);
          }
        }
        else
          ev[i * VPL] = e; // value index i * VPL: after the reduce-scatter lane gl owns the total of sample gl, its own
      }
      if constexpr (!T::kJac)
      {
        // one reduce-scatter of the group's LPG per-sample sums (7 shuffles for 8 samples) instead of LPG group reductions
        float r[VPL];
        reduce_scatter8<LPG>(ev, gl, r);
        eacc = fmaf(fs.w[l], r[0], eacc);
      }
    }

    // ------------------------------------------------------------------ back to lane == sample
    float Gxx = 0.f, Gxy = 0.f, Gyy = 0.f, bx = 0.f, by = 0.f, esum = eacc;
    if constexpr (T::kJac)
    {
      __syncwarp();
      const float *xo = xp + lane * 9;
      const float4 g0 = make_float4(xo[0], xo[1], xo[2], xo[3]);
      const float4 g1 = make_float4(xo[4], xo[5], 0.f, 0.f);
      const float m2 = wm * wm; // g~ and r both carry within_mask (:200, :234)
      Gxx = g0.x * m2; Gxy = g0.y * m2; Gyy = g0.z * m2; bx = g0.w * m2; by = g1.x * m2;
      esum = g1.y;
    }
    if constexpr (!T::kJac && PH_ILV)
    {
      // the reduce-scatter left lane gl of group q with the total of the sample served in step gl: gl * NG + q
      const float wme = __shfl_sync(0xffffffffu, wm, gl * NG + q);
      err_acc += wme * esum;
    }
    else
      err_acc += wm * esum; // feat_error = within_mask * diff^2 (:228)
    inl_acc += valid;

    if constexpr (T::kJac)
    {
      // Cholesky G = L L^T and L rho = b  -> virtual rows y1 = l11 P0 + l21 P1, y2 = l22 P1
      const bool on = valid != 0.f;
      const float l11 = sqrtf(Gxx);
      const float il11 = l11 > 0.f ? 1.0f / l11 : 0.f;
      const float l21 = Gxy * il11;
      const float l22 = sqrtf(fmaxf(Gyy - l21 * l21, 0.f));
      const float il22 = l22 > 0.f ? 1.0f / l22 : 0.f;
      const float rho1 = bx * il11;
      const float rho2 = (by - l21 * rho1) * il22;
      // level-independent projection Jacobian P^ = A^ dp1/dx, A^ = [[1/z,0,-x/z^2],[0,1/z,-y/z^2]] (:241-245 without f)
      const float iz = on ? 1.0f / pz0 : 0.f;
      const float xz = px0 * iz, yz = py0 * iz;
      float P0[7], P1[7];
      if constexpr (T::kMap)
      {
        // p_w = d0 R0 x~ + t0 ; dp1/d(pose0) = R1^T [I | -[p_w]x] and dp1/d(pose1) = -dp1/d(pose0)   (:247-297)
        const float wx = d0 * (fs.R0[0] * hx + fs.R0[1] * hy + fs.R0[2] * hz) + fs.t0[0];
        const float wy = d0 * (fs.R0[3] * hx + fs.R0[4] * hy + fs.R0[5] * hz) + fs.t0[1];
        const float wz = d0 * (fs.R0[6] * hx + fs.R0[7] * hy + fs.R0[8] * hz) + fs.t0[2];
        float a0[3], a1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
          a0[k] = iz * fs.R1[k * 3 + 0] - xz * iz * fs.R1[k * 3 + 2];
          a1[k] = iz * fs.R1[k * 3 + 1] - yz * iz * fs.R1[k * 3 + 2];
        }
        P0[0] = a0[0]; P0[1] = a0[1]; P0[2] = a0[2];
        P0[3] = -a0[1] * wz + a0[2] * wy;
        P0[4] = a0[0] * wz - a0[2] * wx;
        P0[5] = -a0[0] * wy + a0[1] * wx;
        P1[0] = a1[0]; P1[1] = a1[1]; P1[2] = a1[2];
        P1[3] = -a1[1] * wz + a1[2] * wy;
        P1[4] = a1[0] * wz - a1[2] * wx;
        P1[5] = -a1[0] * wy + a1[1] * wx;
      }
      else
      {
        // closed-form Jacobian w.r.t. the left-perturbed relative pose (:680-681)
        P0[0] = iz; P0[1] = 0.f; P0[2] = -xz * iz; P0[3] = -xz * yz; P0[4] = 1.0f + xz * xz; P0[5] = -yz;
        P1[0] = 0.f; P1[1] = iz; P1[2] = -yz * iz; P1[3] = -(1.0f + yz * yz); P1[4] = xz * yz; P1[5] = xz;
      }
      // d pi / d depth (:324-325) without the focal length, and the scale column (:335); scale0 == 0 marks the 6-DoF tracker
      const float jdx = rx * iz - px0 * rz * iz * iz;
      const float jdy = ry * iz - py0 * rz * iz * iz;
      const bool has_scale = T::kMap || fs.scale0 != 0.f;
      P0[6] = has_scale ? jdx * d0 / fs.scale0 : 0.f;
      P1[6] = has_scale ? jdy * d0 / fs.scale0 : 0.f;
      const float c11 = on ? l11 : 0.f, c21 = on ? l21 : 0.f, c22 = on ? l22 : 0.f;
      float v[8], u[8];
#pragma unroll
      for (int k = 0; k < 7; ++k)
      {
        v[k] = c11 * P0[k] + c21 * P1[k];
        u[k] = c22 * P1[k];
      }
      v[7] = on ? rho1 : 0.f;
      u[7] = on ? rho2 : 0.f;
      // code columns: P[:, code_i] = jd * scale0 * basis_i (:331-332)  ->  y1 = k1 c, y2 = k2 c
      const float k1 = (c11 * jdx + c21 * jdy) * fs.scale0;
      const float k2 = (c22 * jdy) * fs.scale0;
      // PH_HALF: stage and fold the two virtual rows of the 32 samples one after the other (half the staging buffer)
#pragma unroll
      for (int pass = 0; pass < (PH_HALF ? 2 : 1); ++pass)
      {
        float *row = PH_HALF ? Yw + (size_t)lane * ST : Yw + (size_t)(2 * lane) * ST;
        if (!PH_HALF || pass == 0)
        {
          *reinterpret_cast<float4 *>(row) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4 *>(row + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (!PH_HALF || pass == 1)
        {
          float *r2 = PH_HALF ? row : row + ST;
          *reinterpret_cast<float4 *>(r2) = make_float4(u[0], u[1], u[2], u[3]);
          *reinterpret_cast<float4 *>(r2 + 4) = make_float4(u[4], u[5], u[6], u[7]);
        }
        if constexpr (T::kMap)
        {
#pragma unroll
          for (int i = 0; i < LPG; ++i) // lane == channel quad
          {
            const int src = q * LPG + i;
            const float s1 = __shfl_sync(0xffffffffu, k1, src), s2 = __shfl_sync(0xffffffffu, k2, src);
            const int sidx = __shfl_sync(0xffffffffu, idx, src);
            float *r0 = (PH_HALF ? Yw + (size_t)src * ST : Yw + (size_t)(2 * src) * ST) + 8;
#pragma unroll
            for (int k = 0; k < CCH; ++k)
            {
              const int ch = gl + k * LPG;
              if (ch < C / 4)
              {
                const float4 cb = ldg4(fs.basis0 + (size_t)sidx * C + ch * 4);
                if (!PH_HALF || pass == 0)
                  *reinterpret_cast<float4 *>(r0 + ch * 4) = make_float4(s1 * cb.x, s1 * cb.y, s1 * cb.z, s1 * cb.w);
                if (!PH_HALF || pass == 1)
                  *reinterpret_cast<float4 *>(r0 + (PH_HALF ? 0 : ST) + ch * 4) = make_float4(s2 * cb.x, s2 * cb.y, s2 * cb.z, s2 * cb.w);
              }
            }
          }
        }
        __syncwarp();
        syrk.accumulate(Yw, ROWS, lane);
        __syncwarp();
      }
    }
    if constexpr (T::kJac)
    {
      if (++nacc == PH_FLUSH) // 3 * ROWS / 8 mma per tile, level and batch since the last flush
      {
        syrk.flush_frag(acc_w, lane);
        nacc = 0;
      }
    }
  }

  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x; // partials are indexed by launch-local factor
  if constexpr (T::kJac)
  {
    if (nacc)
      syrk.flush_frag(acc_w, lane);
    __syncthreads();
    MmaSyrk<WP>::store_frags(acc_s, partH + slot * (WP * WP), PH_WARPS); // upper triangle of the CTA's WP x WP partial
  }
  const float e = block_sum(err_acc, red);
  const float c = block_sum(inl_acc, red);
  if (threadIdx.x == 0)
  {
    partE[slot * 2 + 0] = e;
    partE[slot * 2 + 1] = c;
  }
}

// Reduce the per-CTA partials of one factor in a fixed order and emit the reference outputs
//   [AtA D*D | Atb D | error | inliers]  with AtA = (1/n) sum, error = sum/n, or the zero-overlap fallback
//   (error = 10 * sum w_l, zeros)   photometric_factor_kernels.cpp:1139-1161.
// Internal row layout [pose0 6 | scale | rhs | code C]; the pose1 block of the reference is -pose0.
template <int C, int MODE>
__global__ void photo_finalize_kernel(const PhotoFactor *__restrict__ factors, int nlevels, int slices, const float *__restrict__ partH,
                                      const float *__restrict__ partE, float *__restrict__ out, int out_stride, int D)
{
  constexpr bool kJac = (MODE == PH_MAP_JAC || MODE == PH_TRK_JAC);
  constexpr bool kMap = (MODE == PH_MAP_JAC || MODE == PH_MAP_ERR);
  constexpr int WP = kMap ? 8 + C : 8;
  const PhotoFactor &f = factors[blockIdx.x];
  const int slot = blockIdx.x; // launch-local index into the partials
  float *o = out + (size_t)f.out * out_stride;
  __shared__ float s_n, s_e;
  if (threadIdx.x == 0)
  {
    float e = 0.f, n = 0.f;
    for (int s = 0; s < slices; ++s)
    {
      e += partE[((size_t)slot * slices + s) * 2 + 0];
      n += partE[((size_t)slot * slices + s) * 2 + 1];
    }
    s_n = n;
    float sw = 0.f;
    for (int l = 0; l < nlevels; ++l)
      sw += f.w[l];
    s_e = n > 0.f ? e / n : sw * 10.0f;
  }
  __syncthreads();
  const float n = s_n;
  const int base = kJac ? D * D + D : 0;
  if (threadIdx.x == 0)
  {
    o[base + 0] = s_e;
    o[base + 1] = n;
  }
  if constexpr (kJac)
  {
    const float inv = n > 0.f ? 1.0f / n : 0.f;
    // reference column c -> internal column and sign
    auto icol = [&](int c, float &sg) -> int {
      sg = 1.f;
      if constexpr (kMap)
      {
        if (c < 6)
          return c;
        if (c < 12)
        {
          sg = -1.f;
          return c - 6;
        }
        return c < 12 + C ? 8 + (c - 12) : 6;
      }
      else
        return c;
    };
    for (int e = threadIdx.x; e < D * D + D; e += blockDim.x)
    {
      float sr, sc = 1.f;
      int r, c;
      if (e < D * D)
      {
        r = icol(e / D, sr);
        c = icol(e % D, sc);
      }
      else
      {
        r = icol(e - D * D, sr);
        c = 7;
      }
      const int lo = r < c ? r : c, hi = r < c ? c : r; // only the upper triangle of the partial is written
      float v = 0.f;
      for (int s = 0; s < slices; ++s)
        v += partH[((size_t)slot * slices + s) * (WP * WP) + lo * WP + hi];
      o[e] = sr * sc * v * inv;
    }
  }
}

template <int F, int C, int MODE>
constexpr size_t photo_window_bytes()
{
  return (size_t)((MODE == PH_MAP_JAC || MODE == PH_TRK_JAC) ? PH_WIN_KB_JAC : PH_WIN_KB_ERR) * 1024;
}

// dynamic shared memory: the staged variant's window, then (lineariser modes) the warps' second-level accumulators
template <int F, int C, int MODE, bool STG>
constexpr size_t photo_dyn_bytes()
{
  constexpr bool jac = (MODE == PH_MAP_JAC || MODE == PH_TRK_JAC);
  constexpr int WP = (MODE == PH_MAP_JAC || MODE == PH_MAP_ERR) ? 8 + C : 8;
  return (STG ? photo_window_bytes<F, C, MODE>() : 0) + (jac ? (size_t)PH_WARPS * MmaSyrk<WP>::NTILES * 32 * sizeof(float4) : 0);
}

template <int F, int C, int MODE, bool STG>
static void launch_photo_t(const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH, float *partE,
                           float *out, int out_stride, int D, cudaStream_t stream)
{
  dim3 grid(slices, nfactors);
  constexpr size_t dyn = photo_dyn_bytes<F, C, MODE, STG>();
  if constexpr (dyn != 0)
  {
    static unsigned long long done = 0; // per instantiation and device
    if (first_use_on_device(done))
      cudaFuncSetAttribute(photo_kernel<F, C, MODE, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  }
  photo_kernel<F, C, MODE, STG><<<grid, PH_CTA, dyn, stream>>>(factors, cam, partH, partE);
  photo_finalize_kernel<C, MODE><<<nfactors, 256, 0, stream>>>(factors, cam.L, slices, partH, partE, out, out_stride, D);
}

template <int F, int C>
static void launch_photo_fc(int mode, bool staged, const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH,
                            float *partE, float *out, int out_stride, int D, cudaStream_t stream)
{
  switch (mode)
  {
  case PH_MAP_JAC:
    if (staged)
      launch_photo_t<F, C, PH_MAP_JAC, true>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream);
    else
      launch_photo_t<F, C, PH_MAP_JAC, false>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream);
    break;
  case PH_MAP_ERR:
    if (staged)
      launch_photo_t<F, C, PH_MAP_ERR, true>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream);
    else
      launch_photo_t<F, C, PH_MAP_ERR, false>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream);
    break;
  case PH_TRK_JAC: launch_photo_t<F, C, PH_TRK_JAC, false>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream); break;
  default: launch_photo_t<F, C, PH_TRK_ERR, false>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream); break;
  }
}

int photo_row_width(int mode, int C) { return (mode == PH_MAP_JAC || mode == PH_MAP_ERR) ? 8 + C : 8; }
int photo_samples_per_cta() { return PH_WARPS * 32; }
// floats of one CTA's J^T J partial
size_t photo_partial_floats(int mode, int C)
{
  const size_t WP = photo_row_width(mode, C);
  return (mode == PH_MAP_JAC || mode == PH_TRK_JAC) ? WP * WP : 4;
}

// resident CTAs per SM of the kernel the given configuration launches (occupancy API); 0 for an unsupported (F, C)
int photo_ctas_per_sm(int mode, int F, int C, bool staged)
{
  int n = 0;
#define SAGE_OCC1(FF, CC, MM)                                                                                                   \
  {                                                                                                                             \
    if (staged && (MM == PH_MAP_JAC || MM == PH_MAP_ERR))                                                                       \
    {                                                                                                                           \
      cudaFuncSetAttribute(photo_kernel<FF, CC, MM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,                         \
                           (int)photo_dyn_bytes<FF, CC, MM, true>());                                                           \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, photo_kernel<FF, CC, MM, true>, PH_CTA, photo_dyn_bytes<FF, CC, MM, true>()); \
    }                                                                                                                           \
    else                                                                                                                        \
    {                                                                                                                           \
      if (photo_dyn_bytes<FF, CC, MM, false>() != 0)                                                                            \
        cudaFuncSetAttribute(photo_kernel<FF, CC, MM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                             (int)photo_dyn_bytes<FF, CC, MM, false>());                                                        \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, photo_kernel<FF, CC, MM, false>, PH_CTA, photo_dyn_bytes<FF, CC, MM, false>()); \
    }                                                                                                                           \
  }
#define SAGE_OCC(FF, CC)                                     \
  if (F == FF && C == CC)                                    \
  {                                                          \
    if (mode == PH_MAP_JAC) SAGE_OCC1(FF, CC, PH_MAP_JAC)    \
    else if (mode == PH_MAP_ERR) SAGE_OCC1(FF, CC, PH_MAP_ERR) \
    else if (mode == PH_TRK_JAC) SAGE_OCC1(FF, CC, PH_TRK_JAC) \
    else SAGE_OCC1(FF, CC, PH_TRK_ERR)                       \
  }
  SAGE_OCC(32, 32)
  SAGE_OCC(16, 16)
  SAGE_OCC(16, 8)
  SAGE_OCC(32, 16)
  SAGE_OCC(16, 32)
#undef SAGE_OCC
#undef SAGE_OCC1
  return n;
}

// returns 0 on success, -1 for an unsupported (F, C).  staged: the samples of kf0 are in tile-major order (problem keyframes);
// only the mapping forms have a staged instantiation.
int launch_photo(int mode, int F, int C, const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH,
                 float *partE, float *out, int out_stride, int D, cudaStream_t stream, bool staged)
{
  if (nfactors <= 0)
    return 0;
#define SAGE_CASE(FF, CC)                                                                                               \
  if (F == FF && C == CC)                                                                                               \
  {                                                                                                                     \
    launch_photo_fc<FF, CC>(mode, staged, factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream);   \
    return 0;                                                                                                           \
  }
  SAGE_CASE(32, 32)
  SAGE_CASE(16, 16)
  SAGE_CASE(16, 8)
  SAGE_CASE(32, 16)
  SAGE_CASE(16, 32)
#undef SAGE_CASE
  return -1;
}

} // namespace sage

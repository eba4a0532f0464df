// geometric.cu -- fused geometric (depth-consistency) linearisation / error kernels (sm_100a).
//
// Replaces geometric_jac_error_calculate_kernel (cuda/geometric_factor_kernels.cpp:472-720),
// geometric_error_calculate_kernel (:127-218) and the ATen reductions after them (:931-947, :870-879).
//
// Two linearisers (see launch_geo_c): geo_tc_kernel -- tcgen05.mma with the accumulators in tensor memory, code size 32 -- further
// down, and geo_kernel (mma.sync; also the error-only pass) right here:
// A CTA is a pair of warps; each warp takes batches of 32 samples.
//  * lane == sample: depth of the sample, warp into KF1, nearest mask lookup, the bilinear taps of KF1's
//    (depth, d/dx, d/dy, mask) float4 map, Cauchy weight and the 9 "small" columns of the Jacobian row
//    ([pose0 6 | scale0 | scale1 | rhs]; the pose1 block is exactly -pose0 and is expanded at the end).
//  * lane == channel quad: a group of C/4 lanes fetches the KF0 basis row and the four taps of KF1's
//    pixel-major basis [HW][C] as float4 (one fully used 128-byte line per tap for C = 32) and writes the
//    2C code columns of the row.
//  * The staged rows (width 16 + 2C, one per sample) of both warps are folded into J^T J | J^T r on the
//    tensor cores (mma.sync m16n8k8, 3xTF32, fp32 accumulate), the upper-triangular tiles split between
//    the two warps.  Nothing is written to HBM except one partial per CTA (updated every GEO_FLUSH rounds, see
//    MmaSyrk::flush_tiles for why).
#include "sage_common.cuh"
#include "sage_kernels.h"
#include "tcgen05.cuh"
#include <algorithm>
#include <cstdlib>

namespace sage
{

struct GeoCam
{
  float fx, fy, cx, cy;
  int W, H;
};

#ifndef SAGE_GEO_TC_DEFAULT
#define SAGE_GEO_TC_DEFAULT 1
#endif
#ifndef GEO_TC_FLUSH
#define GEO_TC_FLUSH 8
#endif
#ifndef GEO_FLUSH
#define GEO_FLUSH 8
#endif
#ifndef GEO_MINB
#define GEO_MINB 8
#endif
#ifndef GEO_UNROLL
#define GEO_UNROLL 4
#endif
#define GEO_STR2(x) #x
#define GEO_STR(x) GEO_STR2(x)
constexpr int GEO_WARPS = 2;
constexpr int GEO_CTA = GEO_WARPS * 32;

// staged row: [pose0 6 | scale0 | scale1 | rhs | 0 x7 | code0 C | code1 C]
template <int C>
struct GeoTraits
{
  static constexpr int LPG = C / 4;
  static constexpr int NG = 32 / LPG;
  static constexpr int WP = 16 + 2 * C;
  using Syrk = MmaSyrk<WP, GEO_WARPS>;
  static constexpr int ST = Syrk::ST;
};


// Everything the thread that owns a sample computes once d0 = (bias + basis . code) * scale_0 is known: warp into KF1, nearest
// mask, bilinear taps of KF1's (depth, d/dx, d/dy, mask) map, Cauchy weight and the small columns of the row.
struct GeoSample
{
  float valid, e;  // mask weight (0 = sample does not count), robust error
  float4 c0, c1;   // sw * [pose0 6 | scale0 | scale1]
  float rhs;       // sw * diff
  float k0, k1;    // factors of the code0 / code1 columns
  TapSet tp;       // taps of KF1 (stride 4: pk & ~3 = pixel * 4)
};

// the part of a sample that needs no load from KF1: warp into KF1, nearest-mask pixel, bilinear taps
struct GeoProj
{
  float rx, ry, rz, px, py, pz;
  bool inb;  // the sample is live, in front of KF1 and its nearest pixel lies inside the image
  int near;  // which of the four taps that pixel is: bit 1 = east, bit 0 = south
  TapSet tp; // taps of KF1 (stride 4: pk & ~3 = pixel * 4); at pixel 0 with the zero-padding weights when !inb
};

__device__ __forceinline__ GeoProj geo_project(const GeoFactor &fs, const GeoCam &cam, const float4 hm, const float d0, const bool live)
{
  GeoProj p;
  const int W = cam.W, H = cam.H;
  const float hx = hm.x, hy = hm.y, hz = hm.z;
  p.rx = fs.R10[0] * hx + fs.R10[1] * hy + fs.R10[2] * hz;
  p.ry = fs.R10[3] * hx + fs.R10[4] * hy + fs.R10[5] * hz;
  p.rz = fs.R10[6] * hx + fs.R10[7] * hy + fs.R10[8] * hz;
  p.px = d0 * p.rx + fs.t10[0], p.py = d0 * p.ry + fs.t10[1], p.pz = d0 * p.rz + fs.t10[2];
  const bool pos = p.pz > fs.eps;
  float ux = (p.px / p.pz) * cam.fx + cam.cx;
  float uy = (p.py / p.pz) * cam.fy + cam.cy;
  const int mx = (int)roundf(ux), my = (int)roundf(uy);
  p.inb = live && pos && within(mx, my, W, H);
  // round(u) is floor(u) or floor(u) + 1: the nearest pixel is one of the bilinear taps, and when it lies inside the image that
  // tap's address is not clamped -- the mask (the .w of the map) comes with the taps, no extra load
  p.near = ((mx > (int)floorf(ux)) ? 2 : 0) | ((my > (int)floorf(uy)) ? 1 : 0);
  if (!p.inb)
  {
    ux = 0.f; // keep everything finite; the row is multiplied by the zero weight below
    uy = 0.f;
  }
  p.tp = make_tapset(ux, uy, W, H, 4);
  return p;
}

// wm: mask of KF1 at the nearest pixel (0 when !inb); dgv: bilinear sample of KF1's (depth, d/dx, d/dy, mask) map at the taps
template <bool JAC>
__device__ __forceinline__ GeoSample geo_finish(const GeoFactor &fs, const GeoCam &cam, const float4 hm, const float d0, const GeoProj &p,
                                                const float wm, const float4 dgv)
{
  GeoSample g;
  const float hx = hm.x, hy = hm.y, hz = hm.z;
  const float rx = p.rx, ry = p.ry, rz = p.rz, px = p.px, py = p.py, pz = p.pz;
  g.valid = wm;
  g.tp = p.tp;
  const float D1 = fs.dscale * dgv.x; // sampled (scaled) depth of KF1 and its gradient
  const float gx = fs.dscale * dgv.y;
  const float gy = fs.dscale * dgv.z;
  const float diff = D1 - pz;
  const float md = wm * diff;
  g.e = wm != 0.f ? logf(1.0f + (md * md) / fs.loss_param) : 0.f; // :600
  if constexpr (JAC)
  {
    const bool on = wm != 0.f;
    const float sw = on ? wm * sqrtf(1.0f / (diff * diff + fs.loss_param)) : 0.f; // :690
    const float iz = on ? 1.0f / pz : 0.f;
    const float xz = px * iz, yz = py * iz;
    const float wx = d0 * (fs.R0[0] * hx + fs.R0[1] * hy + fs.R0[2] * hz) + fs.t0[0];
    const float wy = d0 * (fs.R0[3] * hx + fs.R0[4] * hy + fs.R0[5] * hz) + fs.t0[1];
    const float wz = d0 * (fs.R0[6] * hx + fs.R0[7] * hy + fs.R0[8] * hz) + fs.t0[2];
    // v = (R1^T)[2,:] - gx * (A R1^T)[0,:] - gy * (A R1^T)[1,:]  with A the 2x3 projection Jacobian (:607-608, :671-679)
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      const float a0 = cam.fx * iz * fs.R1[k * 3 + 0] - cam.fx * xz * iz * fs.R1[k * 3 + 2];
      const float a1 = cam.fy * iz * fs.R1[k * 3 + 1] - cam.fy * yz * iz * fs.R1[k * 3 + 2];
      v[k] = fs.R1[k * 3 + 2] - (gx * a0 + gy * a1);
    }
    const float p3 = -v[1] * wz + v[2] * wy, p4 = v[0] * wz - v[2] * wx, p5 = -v[0] * wy + v[1] * wx;
    const float jdx = cam.fx * (rx * iz - px * rz * iz * iz);
    const float jdy = cam.fy * (ry * iz - py * rz * iz * iz);
    const float d1_jac_d0 = gx * jdx + gy * jdy;
    const float kc0 = (rz - d1_jac_d0) * fs.scale0;      // :685
    const float js0 = (rz - d1_jac_d0) * d0 / fs.scale0; // :687
    const float js1 = -D1 / fs.scale1;                   // :688
    g.c0 = make_float4(sw * v[0], sw * v[1], sw * v[2], sw * p3);
    g.c1 = make_float4(sw * p4, sw * p5, sw * js0, sw * js1);
    g.rhs = sw * diff;
    g.k0 = sw * kc0;
    g.k1 = -(sw * fs.scale1); // :695-696
  }
  return g;
}

// bilinear sample of KF1's (depth, d/dx, d/dy, mask) map at the taps (the reference's tap order, as gather4) and the mask at the
// nearest pixel (0 when the sample does not project into the image)
__device__ __forceinline__ float4 geo_sample_dgm(const GeoFactor &fs, const GeoCam &cam, const GeoProj &p, float &wm)
{
  const TapSet &tp = p.tp;
  const float *pnw = reinterpret_cast<const float *>(fs.dgm1) + (tp.pk & ~3);
  const float *pne = pnw + ((tp.pk & 2) ? 4 : 0);
  const float *psw = pnw + ((tp.pk & 1) ? 4 * cam.W : 0);
  const float *pse = psw + ((tp.pk & 2) ? 4 : 0);
  const float4 a = ldg4(pnw), b = ldg4(pse), c = ldg4(psw), d = ldg4(pne);
  const float *w = tp.w;
  float4 r;
  r.x = a.x * w[0] + b.x * w[1] + c.x * w[2] + d.x * w[3];
  r.y = a.y * w[0] + b.y * w[1] + c.y * w[2] + d.y * w[3];
  r.z = a.z * w[0] + b.z * w[1] + c.z * w[2] + d.z * w[3];
  r.w = 0.f;
  const float wn = (p.near & 2) ? d.w : a.w, ws = (p.near & 2) ? b.w : c.w;
  wm = p.inb ? ((p.near & 1) ? ws : wn) : 0.f;
  return r;
}

template <bool JAC>
__device__ __forceinline__ GeoSample geo_front(const GeoFactor &fs, const GeoCam &cam, const float4 hm, const float d0, const bool live)
{
  const GeoProj p = geo_project(fs, cam, hm, d0, live);
  float wm;
  const float4 dgv = geo_sample_dgm(fs, cam, p, wm);
  return geo_finish<JAC>(fs, cam, hm, d0, p, wm, dgv);
}

template <int C, bool JAC>
__global__ void __launch_bounds__(GEO_CTA, GEO_MINB)
geo_kernel(const GeoFactor *__restrict__ factors, const GeoCam cam, float *__restrict__ partH, float *__restrict__ partE)
{
  using T = GeoTraits<C>;
  constexpr int LPG = T::LPG, WP = T::WP, ST = T::ST;
  constexpr int ROWS = GEO_WARPS * 32;
  constexpr int STAGE = JAC ? ROWS * ST : 4;
  __shared__ __align__(16) float Y[STAGE];
  __shared__ GeoFactor fs;
  __shared__ float red[32];
  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.y);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(GeoFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % LPG, q = lane / LPG;
  const int N = fs.N;
  const int W = cam.W;
  float *Yw = Y + (size_t)warp * 32 * ST;

  typename T::Syrk syrk;
  if constexpr (JAC)
    syrk.init();
  float err_acc = 0.f, inl_acc = 0.f;
  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  float *part = partH + slot * (WP * WP); // the two warps own disjoint tiles covering the upper triangle
  int nacc = 0;
  bool first = true;

  const int nbatch = (N + 31) / 32;
  const int nround = (nbatch + GEO_WARPS - 1) / GEO_WARPS; // both warps run the same number of rounds (block barriers inside)
  for (int round = blockIdx.x; round < nround; round += gridDim.x)
  {
    const int batch = round * GEO_WARPS + warp;
    // ------------------------------------------------------------------ lane == sample
    const int n = batch * 32 + lane;
    const bool live = n < N;
    const int nc = min(n, N - 1);
    const float4 hm = __ldg(fs.homo + nc);
    const int idx = __ldg(fs.loc1d + nc);
    // dpt_0 = (bias + jac . code) * scale_0   (:515-521); the dot product runs lane == channel quad
    float mydot = 0.f;
#pragma unroll
    for (int i = 0; i < LPG; ++i)
    {
      const int sidx = __shfl_sync(0xffffffffu, idx, q * LPG + i);
      const float4 cb = ldg4(fs.basis0 + (size_t)sidx * C + gl * 4);
      float dot = cb.x * fs.code0[gl * 4 + 0] + cb.y * fs.code0[gl * 4 + 1] + cb.z * fs.code0[gl * 4 + 2] + cb.w * fs.code0[gl * 4 + 3];
      dot = group_sum<LPG>(dot);
      if (gl == i)
        mydot = dot;
    }
    const float d0 = (__ldg(fs.bias0 + idx) + mydot) * fs.scale0;
    const GeoSample g = geo_front<JAC>(fs, cam, hm, d0, live);
    const TapSet &tp = g.tp;
    err_acc += g.e;
    inl_acc += g.valid;

    if constexpr (JAC)
    {
      float *row = Yw + (size_t)lane * ST;
      *reinterpret_cast<float4 *>(row + 0) = g.c0;
      *reinterpret_cast<float4 *>(row + 4) = g.c1;
      *reinterpret_cast<float4 *>(row + 8) = make_float4(g.rhs, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4 *>(row + 12) = make_float4(0.f, 0.f, 0.f, 0.f);
      // ---------------------------------------------------------------- lane == channel quad: code columns
      const float k0 = g.k0, k1 = g.k1;
_Pragma(GEO_STR(unroll GEO_UNROLL))
      for (int i = 0; i < LPG; ++i)
      {
        const int src = q * LPG + i;
        const float s0 = __shfl_sync(0xffffffffu, k0, src), s1 = __shfl_sync(0xffffffffu, k1, src);
        const int sidx = __shfl_sync(0xffffffffu, idx, src);
        const TapSet ts = shfl_tapset(tp, src);
        const float4 c0 = ldg4(fs.basis0 + (size_t)sidx * C + gl * 4);
        const float *pnw = fs.basis1 + (size_t)(ts.pk >> 2) * C + gl * 4;
        const float *pne = pnw + ((ts.pk & 2) ? C : 0);
        const float *psw = pnw + ((ts.pk & 1) ? W * C : 0);
        const float *pse = psw + ((ts.pk & 2) ? C : 0);
        const float4 c1 = gather4(pnw, pse, psw, pne, ts.w);
        float *r = Yw + (size_t)src * ST + 16 + gl * 4;
        *reinterpret_cast<float4 *>(r) = make_float4(s0 * c0.x, s0 * c0.y, s0 * c0.z, s0 * c0.w);
        *reinterpret_cast<float4 *>(r + C) = make_float4(s1 * c1.x, s1 * c1.y, s1 * c1.z, s1 * c1.w);
      }
      __syncthreads();
      if (warp == 0)
        syrk.template accumulate<0>(Y, ROWS, lane);
      else
        syrk.template accumulate<1>(Y, ROWS, lane);
      if (++nacc == GEO_FLUSH) // 24 mma per tile and round: see MmaSyrk::flush_tiles
      {
        if (warp == 0)
          syrk.template flush_tiles<0>(part, lane, first);
        else
          syrk.template flush_tiles<1>(part, lane, first);
        nacc = 0;
        first = false;
      }
      __syncthreads();
    }
  }

  if constexpr (JAC)
  {
    if (nacc || first)
    {
      if (warp == 0)
        syrk.template flush_tiles<0>(part, lane, first);
      else
        syrk.template flush_tiles<1>(part, lane, first);
    }
  }
  const float es = block_sum(err_acc, red);
  const float cs = block_sum(inl_acc, red);
  if (threadIdx.x == 0)
  {
    partE[slot * 2 + 0] = es;
    partE[slot * 2 + 1] = cs;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The lineariser on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory).
//
// A CTA is eight warps; a round is 128 samples = 16 K-steps of 8 rows.  kind::tf32 takes K-major operands only (with the MN-major
// bits set the instruction is a no-op on sm_100a -- sage-slam_b200/csrc/tc_probe.cu, test 1), so rows are staged K-major without
// swizzle: a core matrix is 8 columns x 4 consecutive samples (8 x 16 bytes = 128 contiguous bytes).  Element (operand column
// m', sample kk of K-step s) lives at
//     s * KSTEP + (m' / 8) * SBO + (kk / 4) * LBO + (m' % 8) * 16 + (kk % 4) * 4,      SBO = 272, LBO = 144, KSTEP = 2 WP / 8 * SBO
// (the strides are padded from 256 / 128 so that the 32 four-byte stores of a warp fall into 32 different banks, see below).
// The operand columns of a K-step are ordered  m' = [ loB | hi | loA ]  with hi = the WP tf32 "hi" halves of the row, loA = the
// "lo" halves of its first L1 = 128 - WP columns and loB = those of the other L2 = WP - L1.  ONE instruction per K-step,
//   A = [hi | loA] (M = 128, starts at column L2),   B = [loB | hi] (N = L2 + WP = 112, starts at column 0),
// leaves in the accumulator  D[0:WP, 0:L2] = hi^T loB,  D[0:WP, L2:] = hi^T hi  and  D[WP:128, L2:] = loA^T hi.  Since A == B the
// third product of the 3xTF32 split, lo^T hi, is the transpose of X = hi^T lo, so all of X is there -- its columns >= L1 directly,
// its columns < L1 as the transposed rows WP.. -- and
//   J^T J | J^T r  =  hi^T hi + X + X^T,   formed by the finalize kernel from the per-CTA partial [128][NCOL].
// Thread <-> data for the code columns: a quarter-warp is one sample x the eight channel quads (one full 128-byte line of the
// basis per load: the LSU handles a 16-byte-per-lane load one quarter-warp at a time, so eight lanes on eight different lines
// cost eight passes of the L1 data pipe instead of one -- measured: 28 instead of 5 wavefronts per request), the four quarters
// are four CONSECUTIVE samples (one 16-byte K-chunk).  A thread then holds 4 consecutive columns of one sample and writes them
// as four 4-byte stores 16 bytes apart; bank = (sample % 4) + 4 (quad / 2) + 16 (quad % 2): conflict-free because SBO = 272
// bytes = 68 words = 4 mod 32.  The single issuing thread never waits for the tensor pipe inside a round: the stage is handed
// over with tcgen05.commit -> mbarrier, and the other CTA of the SM gathers while this one multiplies.
// The stage costs 680 bytes per sample, so only 256 samples are in flight per SM; to have 16 warps on them a warp takes 16
// samples per round (half the lanes idle while lane == sample, all of them busy on the code columns), and the sample indices
// of the next round are fetched one round ahead.
constexpr int GTC_THREADS = 256; // 8 warps x 16 samples = one round of 128 samples
constexpr int GTC_ROUND = 128;

template <int C>
struct GeoTc
{
  static_assert(C == 32, "a quarter-warp covers the eight channel quads of one sample");
  static constexpr int WP = 16 + 2 * C;
  static constexpr int NG8 = 2 * WP / 8;       // 8-column groups of the operand (hi + lo)
  static constexpr int SBO = 272, LBO = 144;   // bytes between 8-column groups / between the two 4-sample halves of a K-step
  static constexpr int KSTEP = NG8 * SBO;      // (KSTEP / 4) % 32 == 16: the two K-steps a warp's own samples go to use different banks
  static constexpr int KSTEPS = GTC_ROUND / 8;
  static constexpr int STAGE = KSTEPS * KSTEP + (NG8 < 16 ? (16 - NG8) * SBO : 0); // the A operand reads 16 groups of every step
  static constexpr int L1 = 128 - WP;                      // lo columns inside the 128 rows of A (loA)
  static constexpr int L2 = WP - L1;                       // the other lo columns (loB), part of B
  static constexpr int NCOL = WP + L2;                     // accumulator columns = N of the instruction
  static constexpr int HI0 = L2 / 8, LOA0 = (L2 + WP) / 8; // first 8-column group of hi / loA (loB starts at 0)
  static constexpr int TCOLS = NCOL > 128 ? 256 : (NCOL > 64 ? 128 : 64);
  static constexpr int LO = (WP / 8) * SBO;    // byte offset from a hi element to its lo element: + for columns < L1, - for the others
  static constexpr int PART = 128 * NCOL;      // floats of one CTA's partial
  static_assert(WP > 64 && L1 % 8 == 0 && L2 % 8 == 0 && NCOL % 16 == 0 && NCOL <= 256, "tcgen05.mma shapes");
  static_assert(16 + C <= L1 && L1 == 16 + C, "[small | code0] are the loA columns, code1 the loB columns");
  static_assert((KSTEP / 4) % 32 == 16 && (SBO / 4) % 32 == 4 && LBO + 128 <= SBO && KSTEP % 16 == 0, "bank rotation / descriptor alignment");
};

// four consecutive columns of one sample: p = address of (first column, sample), the next columns are 16 bytes apart
template <int LO>
__device__ __forceinline__ void st_hi_lo(unsigned char *p, const float4 v)
{
  uint32_t h, l;
  split_tf32(v.x, h, l);
  *reinterpret_cast<uint32_t *>(p) = h;
  *reinterpret_cast<uint32_t *>(p + LO) = l;
  split_tf32(v.y, h, l);
  *reinterpret_cast<uint32_t *>(p + 16) = h;
  *reinterpret_cast<uint32_t *>(p + 16 + LO) = l;
  split_tf32(v.z, h, l);
  *reinterpret_cast<uint32_t *>(p + 32) = h;
  *reinterpret_cast<uint32_t *>(p + 32 + LO) = l;
  split_tf32(v.w, h, l);
  *reinterpret_cast<uint32_t *>(p + 48) = h;
  *reinterpret_cast<uint32_t *>(p + 48 + LO) = l;
}

__device__ __forceinline__ void red_add4(float *p, const float4 v)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int C>
__global__ void __launch_bounds__(GTC_THREADS, 2)
geo_tc_kernel(const GeoFactor *__restrict__ factors, const GeoCam cam, float *__restrict__ partH, float *__restrict__ partE)
{
  using T = GeoTc<C>;
  constexpr int KSTEP = T::KSTEP, LO = T::LO, NCOL = T::NCOL;
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ GeoFactor fs;
  __shared__ float red[32];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.y);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(GeoFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0)
    tc::tmem_alloc(&tmem_slot, T::TCOLS);
  if (threadIdx.x == 0)
  {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  // columns 9..15 (hi and lo) of every sample stay zero for the whole kernel; nothing else ever writes them
  for (int e = threadIdx.x; e < T::KSTEPS * 8 * 7; e += blockDim.x)
  {
    const int ks = e / 56, kk = (e % 56) / 7, m = 9 + e % 7;
    unsigned char *z = stage + (size_t)ks * KSTEP + (T::HI0 + (m >> 3)) * T::SBO + (kk >> 2) * T::LBO + (m & 7) * 16 + (kk & 3) * 4;
    *reinterpret_cast<uint32_t *>(z) = 0u;
    *reinterpret_cast<uint32_t *>(z + LO) = 0u;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  // lane = q * 8 + gl.  Code columns: step i (0..3) of a round handles the warp's samples j = 4 i + q, lane (q, gl) the channel
  // quad gl of it.  lane == sample: the lanes with gl < 4 own sample j = 4 gl + q (so the owner of step i's sample sits in the
  // same quarter, at gl == i); the lanes with gl >= 4 shadow them and are masked out.
  const int gl = lane & 7, q = lane >> 3;
  const int jown = 4 * (gl & 3) + q;
  const int N = fs.N;
  const int W = cam.W;
  float err_acc = 0.f, inl_acc = 0.f;
  uint32_t phase = 0;
  int pending = 0;  // a commit is outstanding: the tensor core may still be reading the stage
  int nacc = 0;     // rounds accumulated in tensor memory since the last flush
  bool have_partial = false;
  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  float *part = partH + slot * (size_t)T::PART;
  // The tensor core adds into the fp32 accumulator with truncation: every instruction loses half an ulp of the running sum on
  // average, a bias of -6e-8 per instruction that does not average out (measured: 10240 chained instructions -> -5.8e-4 on the
  // diagonal of J^T J).  So the chain is cut every GEO_TC_FLUSH rounds (x 16 K-steps): the accumulator is added to the CTA's
  // partial in global memory with round-to-nearest fp32 adds -- row m belongs to one thread, which adds with fire-and-forget
  // vector reductions (no read latency; one thread's reductions to one address apply in program order, so the sum is
  // deterministic) -- and the next instruction starts from zero.  Accumulator row m = tensor-memory lane m: a warp reads the
  // lanes [32 (w % 4), 32 (w % 4) + 32) (the hardware's rule), warps 0-3 take the even 16-column blocks, warps 4-7 the odd ones.
  auto flush = [&]() {
    tc::fence_after_sync();
    const int wq = warp & 3;
    float *row = part + (size_t)(wq * 32 + lane) * NCOL;
#pragma unroll 1
    for (int c = (warp >> 2) * 16; c < NCOL; c += 32)
    {
      float v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + c, v);
#pragma unroll
      for (int i = 0; i < 16; i += 4)
      {
        const float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        if (have_partial)
          red_add4(row + c + i, o);
        else
          *reinterpret_cast<float4 *>(row + c + i) = o;
      }
    }
    tc::fence_before_sync(); // the loads above are ordered before the instructions issued after the next block barrier
    have_partial = true;
    nacc = 0;
  };

  const int nround = (N + GTC_ROUND - 1) / GTC_ROUND;
  const int own = warp * 16 + jown;
  auto clampn = [&](int round) { return min(round * GTC_ROUND + own, N - 1); };
  int round = blockIdx.x;
  int idx = __ldg(fs.loc1d + clampn(round));
  float4 hm = __ldg(fs.homo + clampn(round));

  for (; round < nround; round += gridDim.x)
  {
    const bool live = gl < 4 && round * GTC_ROUND + own < N;
    // sample indices of the next round
    const int idx_n = __ldg(fs.loc1d + clampn(round + gridDim.x));
    const float4 hm_n = __ldg(fs.homo + clampn(round + gridDim.x));
    // ------------------------------------------------------------------ dpt_0 = (bias + jac . code) * scale_0 (:515-521)
    const float bias = __ldg(fs.bias0 + idx);
    float4 cb[4]; // step i -> sample 4 i + q, channels 4 gl .. + 3; kept for the code columns
    float mydot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const int sidx = __shfl_sync(0xffffffffu, idx, q * 8 + i);
      cb[i] = ldg4(fs.basis0 + (size_t)sidx * C + gl * 4);
      float dot = cb[i].x * fs.code0[gl * 4 + 0] + cb[i].y * fs.code0[gl * 4 + 1] + cb[i].z * fs.code0[gl * 4 + 2] + cb[i].w * fs.code0[gl * 4 + 3];
      dot = group_sum<8>(dot);
      if (gl == i)
        mydot = dot;
    }
    const float d0 = (bias + mydot) * fs.scale0;
    const GeoProj pj = geo_project(fs, cam, hm, d0, live);
    // ------------------------------------------------------------------ everything this round reads from KF1
    float wm;
    const float4 dgv = geo_sample_dgm(fs, cam, pj, wm);
    float4 c1[4]; // bilinear sample of KF1's basis at the taps of sample 4 i + q, channel quad gl
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const TapSet ts = shfl_tapset(pj.tp, q * 8 + i);
      const float *pnw = fs.basis1 + (size_t)(ts.pk >> 2) * C + gl * 4;
      const float *pne = pnw + ((ts.pk & 2) ? C : 0);
      const float *psw = pnw + ((ts.pk & 1) ? W * C : 0);
      const float *pse = psw + ((ts.pk & 2) ? C : 0);
      c1[i] = gather4(pnw, pse, psw, pne, ts.w);
    }
    const GeoSample g = geo_finish<true>(fs, cam, hm, d0, pj, wm, dgv);
    err_acc += g.e; // 0 for the shadow lanes (not live -> zero weight)
    inl_acc += g.valid;

    // the previous round's instructions must be done reading the stage before it is overwritten
    if (pending)
    {
      tc::mbar_wait(&bar, phase);
      phase ^= 1;
      pending = 0;
    }
    if (nacc == GEO_TC_FLUSH)
      flush();
    unsigned char *grp = stage + (size_t)(warp * 2) * KSTEP; // the warp's two K-steps
    if (gl < 4)
    {
      // own sample j = 4 gl + q: K-step j / 8 = gl / 2, sample (j % 8) = 4 (gl % 2) + q of it
      unsigned char *row = grp + (size_t)(gl >> 1) * KSTEP + (gl & 1) * T::LBO + q * 4 + T::HI0 * T::SBO;
      st_hi_lo<LO>(row, g.c0);      // columns 0..3
      st_hi_lo<LO>(row + 64, g.c1); // 4..7
      uint32_t h, l;
      split_tf32(g.rhs, h, l); // 8
      *reinterpret_cast<uint32_t *>(row + T::SBO) = h;
      *reinterpret_cast<uint32_t *>(row + T::SBO + LO) = l;
    }
    // ------------------------------------------------------------------ code columns of sample 4 i + q, channel quad gl
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const int src = q * 8 + i;
      const float s0 = __shfl_sync(0xffffffffu, g.k0, src), s1 = __shfl_sync(0xffffffffu, g.k1, src);
      const float4 a = cb[i], b = c1[i];
      // K-step i / 2, sample 4 (i % 2) + q of it; columns 16 + 4 gl .. + 3 (code0) and C further (code1): 8-column group
      // 2 + gl / 2, column 4 (gl % 2) of it
      unsigned char *dst = grp + (size_t)(i >> 1) * KSTEP + (i & 1) * T::LBO + q * 4 + (T::HI0 + 2 + (gl >> 1)) * T::SBO + (gl & 1) * 64;
      st_hi_lo<LO>(dst, make_float4(s0 * a.x, s0 * a.y, s0 * a.z, s0 * a.w));                        // lo halves: loA, after hi
      st_hi_lo<-LO>(dst + (C / 8) * T::SBO, make_float4(s1 * b.x, s1 * b.y, s1 * b.z, s1 * b.w)); // lo halves: loB, before hi
    }
    tc::fence_proxy_async(); // generic-proxy stores -> operand reads of the tensor core
    __syncthreads();
    if (threadIdx.x == 0)
    {
      tc::fence_after_sync();
      constexpr uint32_t idesc = tc::idesc_tf32(128, NCOL, false);
      // the descriptors of K-step k differ from those of step 0 in the start-address field only (16-byte units, no carry: < 256 KB)
      const uint64_t db0 = tc::smem_desc(tc::smem_u32(stage), T::SBO, T::LBO); // B = [loB | hi]
      const uint64_t da0 = db0 + (uint64_t)((T::HI0 * T::SBO) >> 4);            // A = [hi | loA]
      const uint32_t acc0 = nacc != 0;
#pragma unroll
      for (int k = 0; k < T::KSTEPS; ++k)
        tc::mma_tf32_ss(tmem, da0 + (uint64_t)((k * KSTEP) >> 4), db0 + (uint64_t)((k * KSTEP) >> 4), idesc, k ? 1u : acc0);
      tc::mma_commit(&bar);
    }
    pending = 1;
    ++nacc;
    idx = idx_n, hm = hm_n;
  }

  if (pending)
    tc::mbar_wait(&bar, phase);
  if (nacc)
    flush();
  if (!have_partial) // a CTA without a round still owns a partial
    for (int e = threadIdx.x; e < T::PART; e += blockDim.x)
      part[e] = 0.f;
  tc::fence_before_sync();
  const float es = block_sum(err_acc, red);
  const float cs = block_sum(inl_acc, red);
  if (threadIdx.x == 0)
  {
    partE[slot * 2 + 0] = es;
    partE[slot * 2 + 1] = cs;
  }
  __syncthreads();
  if (warp == 0)
    tc::tmem_dealloc(tmem, T::TCOLS);
}

// [AtA D*D | Atb D | error | inliers], D = 14 + 2C, scaled by weight / n_inliers; zero overlap -> 10*weight, zeros (:931-947)
// TCP: partials come from geo_tc_kernel (the [128][NCOL] accumulator image, see there) instead of geo_kernel ([WP][WP], upper triangle)
template <int C, bool JAC, bool TCP = false>
__global__ void geo_finalize_kernel(const GeoFactor *__restrict__ factors, int slices, const float *__restrict__ partH,
                                    const float *__restrict__ partE, float *__restrict__ out, int out_stride)
{
  constexpr int WP = 16 + 2 * C;
  constexpr int D = 14 + 2 * C;
  const GeoFactor &f = factors[blockIdx.x];
  const int slot = blockIdx.x;
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
    s_e = n > 0.f ? f.weight / n * e : f.weight * 10.0f;
  }
  __syncthreads();
  const float n = s_n;
  const int base = JAC ? D * D + D : 0;
  if (threadIdx.x == 0 && blockIdx.y == 0)
  {
    o[base + 0] = s_e;
    o[base + 1] = n;
  }
  if constexpr (JAC)
  {
    const float sc = n > 0.f ? f.weight / n : 0.f;
    // reference order [pose0 6 | pose1 6 | code0 C | code1 C | scale0 | scale1] -> internal column and sign
    auto icol = [](int c, float &sg) -> int {
      sg = 1.f;
      if (c < 6)
        return c;
      if (c < 12)
      {
        sg = -1.f;
        return c - 6;
      }
      return c < 12 + 2 * C ? 16 + (c - 12) : 6 + (c - 12 - 2 * C);
    };
    // grid.y CTAs share the elements of one factor (the slice partials are 180 x slices x WP^2 floats: keep the machine busy)
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < D * D + D; e += blockDim.x * gridDim.y)
    {
      float sr, scn = 1.f;
      int r, c;
      if (e < D * D)
      {
        r = icol(e / D, sr);
        c = icol(e % D, scn);
      }
      else
      {
        r = icol(e - D * D, sr);
        c = 8;
      }
      const int lo = r < c ? r : c, hi = r < c ? c : r;
      float v = 0.f;
      if constexpr (TCP)
      {
        using T = GeoTc<32>;
        constexpr int NC = T::NCOL, L1 = T::L1, L2 = T::L2;
        // X^T[a][b] = (lo^T hi)[a][b]: accumulator row WP + a for a < L1, else the transpose of the hi^T loB block
        auto xt = [&](const float *P, int a, int b) { return a < L1 ? P[(WP + a) * NC + L2 + b] : P[b * NC + (a - L1)]; };
        for (int s = 0; s < slices; ++s)
        {
          const float *P = partH + ((size_t)slot * slices + s) * T::PART;
          v += P[lo * NC + L2 + hi] + (xt(P, lo, hi) + xt(P, hi, lo));
        }
      }
      else
      {
        for (int s = 0; s < slices; ++s)
          v += partH[((size_t)slot * slices + s) * (WP * WP) + lo * WP + hi];
      }
      o[e] = sr * scn * v * sc;
    }
  }
}

int geo_row_width(int C) { return 16 + 2 * C; }

// which lineariser: tcgen05 (geo_tc_kernel, C % 16 == 0) or mma.sync (geo_kernel).  Process-wide; SAGE_BA_GEO_TC=0/1 sets the
// initial value, sage_ba_set_geometric_tcgen05 changes it (before a problem is created: slice counts and partial buffers depend on it)
static int &geo_tc_flag()
{
  static int on = [] {
    const char *e = getenv("SAGE_BA_GEO_TC");
    return e ? (atoi(e) != 0 ? 1 : 0) : SAGE_GEO_TC_DEFAULT;
  }();
  return on;
}
int geo_set_tc(int on)
{
  const int prev = geo_tc_flag();
  if (on >= 0)
    geo_tc_flag() = on ? 1 : 0;
  return prev;
}
static bool geo_tc_enabled() { return geo_tc_flag() != 0; }
bool geo_uses_tc(bool jac, int C) { return jac && C == 32 && geo_tc_enabled(); }
// floats of one CTA's partial
size_t geo_partial_floats(bool jac, int C, bool tc)
{
  const size_t WP = 16 + 2 * (size_t)C;
  if (!jac)
    return 4;
  if (tc && C == 32)
    return GeoTc<32>::PART;
  return WP * WP;
}

template <int C>
static void launch_geo_tc(const GeoFactor *factors, int nfactors, const GeoCam &cam, int slices, float *partH, float *partE, float *out,
                          int out_stride, cudaStream_t stream)
{
  if constexpr (C == 32)
  {
    static unsigned long long done = 0;
    if (first_use_on_device(done))
      cudaFuncSetAttribute(geo_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, GeoTc<C>::STAGE);
    geo_tc_kernel<C><<<dim3(slices, nfactors), GTC_THREADS, GeoTc<C>::STAGE, stream>>>(factors, cam, partH, partE);
    geo_finalize_kernel<C, true, true><<<dim3(nfactors, 6), 256, 0, stream>>>(factors, slices, partH, partE, out, out_stride);
  }
}

template <int C>
static void launch_geo_c(bool jac, bool tc, const GeoFactor *factors, int nfactors, const GeoCam &cam, int slices, float *partH,
                         float *partE, float *out, int out_stride, cudaStream_t stream)
{
  dim3 grid(slices, nfactors);
  if (jac && tc && C == 32)
    launch_geo_tc<C>(factors, nfactors, cam, slices, partH, partE, out, out_stride, stream);
  else if (jac)
  {
    geo_kernel<C, true><<<grid, GEO_CTA, 0, stream>>>(factors, cam, partH, partE);
    geo_finalize_kernel<C, true><<<dim3(nfactors, 6), 256, 0, stream>>>(factors, slices, partH, partE, out, out_stride);
  }
  else
  {
    geo_kernel<C, false><<<grid, GEO_CTA, 0, stream>>>(factors, cam, partH, partE);
    geo_finalize_kernel<C, false><<<nfactors, 256, 0, stream>>>(factors, slices, partH, partE, out, out_stride);
  }
}

// resident CTAs per SM of the kernel the given configuration launches (occupancy API); 0 for an unsupported C
int geo_ctas_per_sm(bool jac, int C, bool tc)
{
  int n = 0;
  if (jac && tc && C == 32)
  {
    // registers, shared memory (stage + static + 1 KB the driver reserves per CTA) and tensor-memory columns (512 per SM)
    cudaFuncAttributes fa{};
    size_t stage = 0;
    int tcols = 0;
    cudaFuncGetAttributes(&fa, geo_tc_kernel<32>), stage = GeoTc<32>::STAGE, tcols = GeoTc<32>::TCOLS;
    int dev = 0, smem_sm = 0, regs_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    const int by_regs = regs_sm / std::max(1, ((fa.numRegs + 7) / 8 * 8) * GTC_THREADS);
    const int by_smem = (int)(smem_sm / (stage + fa.sharedSizeBytes + 1024));
    return std::max(1, std::min(std::min(by_regs, by_smem), 512 / tcols));
  }
#define SAGE_OCC(CC)                                                                                          \
  if (C == CC)                                                                                                \
    jac ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, geo_kernel<CC, true>, GEO_CTA, 0)                 \
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, geo_kernel<CC, false>, GEO_CTA, 0);
  SAGE_OCC(32)
  SAGE_OCC(16)
  SAGE_OCC(8)
#undef SAGE_OCC
  return n;
}

int launch_geo(bool jac, int C, const GeoFactor *factors, int nfactors, int W, int H, float fx, float fy, float cx, float cy,
               int slices, float *partH, float *partE, float *out, int out_stride, cudaStream_t stream, bool tc)
{
  if (nfactors <= 0)
    return 0;
  GeoCam cam{fx, fy, cx, cy, W, H};
  switch (C)
  {
  case 32: launch_geo_c<32>(jac, tc, factors, nfactors, cam, slices, partH, partE, out, out_stride, stream); return 0;
  case 16: launch_geo_c<16>(jac, tc, factors, nfactors, cam, slices, partH, partE, out, out_stride, stream); return 0;
  case 8: launch_geo_c<8>(jac, tc, factors, nfactors, cam, slices, partH, partE, out, out_stride, stream); return 0;
  default: return -1;
  }
}

} // namespace sage

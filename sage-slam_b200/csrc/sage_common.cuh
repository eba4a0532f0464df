// sage_common.cuh -- device-side building blocks shared by the factor kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SAGE_MAX_LEVELS 8
#define SAGE_MAX_CODE 32
#define SAGE_CTA 256

// CameraPyramid<float> (common/camera_pyramid.h:18-32) passed by value to every kernel instead of the
// reference's per-call thrust::device_vector fill (photometric_factor_kernels.cpp:1099-1106).
struct CamPyr
{
  int L;
  float fx[SAGE_MAX_LEVELS], fy[SAGE_MAX_LEVELS];
  int w[SAGE_MAX_LEVELS], h[SAGE_MAX_LEVELS];
  int off[SAGE_MAX_LEVELS];
  float ofx, ofy, ocx, ocy; // level-0 intrinsics
  int ow, oh;
};

// One photometric-type factor (mapping or tracker form).  State-dependent fields are filled on the
// host (single-factor API) or by setup kernels (batched problem).
struct PhotoFactor
{
  const float *fg0;    // KF0 channel-last [SP][3][F] (mapping form) or nullptr
  const float *fg1;    // frame-1 channel-last [SP][3][F]
  const float *mask1;  // [H][W]
  const float *bias0;  // [HW]
  const float *basis0; // [HW][C] pixel-major
  const int *loc1d;    // [N]
  const float4 *homo;  // [N] (hx, hy, hz, float(idx % W)... unused w)
  const float *sfeat0; // tracker form: [L][N][F]
  const float *dpts0;  // tracker form: [N]
  int N;
  float R10[9], t10[3], R0[9], t0[3], R1[9];
  float code0[SAGE_MAX_CODE];
  float scale0, eps;
  float w[SAGE_MAX_LEVELS];
  int out; // slot in the factor output buffers
  float dmul; // tracker form: depth = dmul * dpts0[n] (TrackFrame optimises the scale of frame 0)
};

struct GeoFactor
{
  const float *bias0, *basis0; // KF0
  const int *loc1d;
  const float4 *homo;
  const float4 *dgm1;  // KF1 [HW] (depth, d/dx, d/dy, mask)  -- depth/gradient UNSCALED when dscale != 1
  const float *basis1; // KF1 [HW][C]
  int N;
  float R10[9], t10[3], R0[9], t0[3], R1[9];
  float code0[SAGE_MAX_CODE];
  float scale0, scale1, dscale, eps, loss_param, weight;
  int out;
};

struct ReprojFactor
{
  const float *bias0, *basis0; // KF0 (mapping form) or nullptr
  const int *loc1d;            // [M]
  const float *homo;           // [M][3]
  const float *match2d;        // [M][2]
  const float *dpts0;          // tracker form [M]
  int M;
  float R10[9], t10[3], R0[9], t0[3], R1[9];
  float code0[SAGE_MAX_CODE];
  float scale0, eps, loss_param, weight;
  float fx, fy, cx, cy;
  int out;
};

// tracker match-geometry factor (3-D point-to-point, Fair loss): K/match_geometry_factor_kernels.cpp:83-292
struct MatchGeomFactor
{
  const float *dpts0, *dpts1; // [M] depth of the keypoint in frame 0 (before dmul) and of its match in frame 1
  const float *homo0, *homo1; // [M][3]
  int M;
  float R[9], t[3];
  float dmul, scale0; // depth multiplier; scale0 == 0 -> 6-DoF form (no scale column)
  float loss_param, weight;
  int out;
};

// mapping-side match-geometry factor (K/match_geometry_factor_kernels.cpp:421-1039, :1090-1359) and its loop-closure
// form (:296-417, :1043-1086; bias0 == nullptr, unscaled depths given per match).
struct MapMatchGeomFactor
{
  const float *bias0, *basis0, *bias1, *basis1; // [HW], [HW][C] pixel-major ; nullptr -> loop form
  const int *loc0, *loc1;                       // [M]
  const float *dpts0, *dpts1;                   // loop form: [M] unscaled depths
  const float *homo0, *homo1;                   // [M][3]
  int M;
  float R10[9], t10[3], R0[9], t0[3], R1[9];
  float code0[SAGE_MAX_CODE], code1[SAGE_MAX_CODE];
  float scale0, scale1, loss_param, weight;
  int loss_type; // 0 fair, 1 L2, 2 huber, 3 unbiased
  int out;
};

__device__ __forceinline__ bool within(int x, int y, int W, int H) { return x >= 0 && x < W && y >= 0 && y < H; }

// bilinear tap set: floor / floor+1, weights from those integers (photometric_factor_kernels.cpp:146-156)
struct Taps
{
  int x0, y0;
  float nw, se, sw, ne;
  bool bnw, bse, bsw, bne;
};

__device__ __forceinline__ Taps make_taps(float px, float py, int W, int H)
{
  Taps t;
  const float fxf = floorf(px), fyf = floorf(py);
  t.x0 = (int)fxf;
  t.y0 = (int)fyf;
  const float lx = (float)(t.x0 + 1) - px, ly = (float)(t.y0 + 1) - py;
  const float ux = 1.0f - lx, uy = 1.0f - ly;
  t.nw = lx * ly;
  t.se = ux * uy;
  t.sw = lx * uy;
  t.ne = ux * ly;
  t.bnw = within(t.x0, t.y0, W, H);
  t.bse = within(t.x0 + 1, t.y0 + 1, W, H);
  t.bsw = within(t.x0, t.y0 + 1, W, H);
  t.bne = within(t.x0 + 1, t.y0, W, H);
  return t;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// acc = nw*a + se*b + sw*c + ne*d in the reference's tap order (nw + se + sw + ne), per component
__device__ __forceinline__ float4 tap_combine(const Taps &t, float4 a, float4 b, float4 c, float4 d)
{
  float4 r;
  r.x = a.x * t.nw + b.x * t.se + c.x * t.sw + d.x * t.ne;
  r.y = a.y * t.nw + b.y * t.se + c.y * t.sw + d.y * t.ne;
  r.z = a.z * t.nw + b.z * t.se + c.z * t.sw + d.z * t.ne;
  r.w = a.w * t.nw + b.w * t.se + c.w * t.sw + d.w * t.ne;
  return r;
}

// reduce over the lanes of a sub-warp group of LPG lanes (power of two), result in every lane
template <int LPG>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }

// block-wide sum of one float per thread; result valid in thread 0. red: >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float *red)
{
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0)
    red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0)
  {
    r = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// packed bilinear tap descriptor shared between lanes (photometric / geometric gathers)
struct TapSet
{
  int pk;      // pixel_offset * stride | (dx << 1) | dy : clamped base element and whether the +1 taps move
  float w[4];  // nw, se, sw, ne ; 0 for out-of-bounds taps (zero padding)
};

// stride = floats per pixel of the map (a multiple of 4, so the two low bits are free for the flags)
__device__ __forceinline__ TapSet make_tapset(float px, float py, int W, int H, int stride)
{
  TapSet t;
  const int x0 = (int)floorf(px), y0 = (int)floorf(py);
  const float lx = (float)(x0 + 1) - px, ly = (float)(y0 + 1) - py;
  const float ux = 1.0f - lx, uy = 1.0f - ly;
  const int x1 = x0 < 0x7fffffff ? x0 + 1 : x0, y1 = y0 < 0x7fffffff ? y0 + 1 : y0; // cvt saturates; avoid wrap-around
  const bool bx0 = x0 >= 0 && x0 < W, bx1 = x1 >= 0 && x1 < W;
  const bool by0 = y0 >= 0 && y0 < H, by1 = y1 >= 0 && y1 < H;
  t.w[0] = (bx0 && by0) ? lx * ly : 0.f; // nw
  t.w[1] = (bx1 && by1) ? ux * uy : 0.f; // se
  t.w[2] = (bx0 && by1) ? lx * uy : 0.f; // sw
  t.w[3] = (bx1 && by0) ? ux * ly : 0.f; // ne
  const int xa = min(max(x0, 0), W - 1), xb = min(max(x1, 0), W - 1);
  const int ya = min(max(y0, 0), H - 1), yb = min(max(y1, 0), H - 1);
  t.pk = ((ya * W + xa) * stride) | ((xb != xa) ? 2 : 0) | ((yb != ya) ? 1 : 0);
  return t;
}

__device__ __forceinline__ TapSet shfl_tapset(const TapSet &t, int src)
{
  TapSet r;
  r.pk = __shfl_sync(0xffffffffu, t.pk, src);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    r.w[k] = __shfl_sync(0xffffffffu, t.w[k], src);
  return r;
}

// 4-tap weighted sum of one float4 map in the reference's tap order (nw + se + sw + ne).  GENERIC: the pointers may address
// shared memory (staged windows) or global memory -> generic loads instead of ld.global.nc
template <bool GENERIC = false>
__device__ __forceinline__ float4 gather4(const float *pnw, const float *pse, const float *psw, const float *pne, const float *w)
{
  float4 a, b, c, d;
  if constexpr (GENERIC)
  {
    a = *reinterpret_cast<const float4 *>(pnw);
    b = *reinterpret_cast<const float4 *>(pse);
    c = *reinterpret_cast<const float4 *>(psw);
    d = *reinterpret_cast<const float4 *>(pne);
  }
  else
  {
    a = ldg4(pnw);
    b = ldg4(pse);
    c = ldg4(psw);
    d = ldg4(pne);
  }
  float4 r;
  r.x = a.x * w[0] + b.x * w[1] + c.x * w[2] + d.x * w[3];
  r.y = a.y * w[0] + b.y * w[1] + c.y * w[2] + d.y * w[3];
  r.z = a.z * w[0] + b.z * w[1] + c.z * w[2] + d.z * w[3];
  r.w = a.w * w[0] + b.w * w[1] + c.w * w[2] + d.w * w[3];
  return r;
}


// ------------------------------------------------------------------------------------------------
// Cooperative symmetric rank-k accumulator: H += Y^T Y for rows Y[r][0..WP) staged in shared memory.
// The WP x WP result is kept as upper-triangular 4x4 register tiles, one tile per thread, with the
// rows split round-robin over KS thread groups (K-split) so that KS * NT threads are busy.
//   NB = WP/4 tile rows, NT = NB*(NB+1)/2 tiles.
// ------------------------------------------------------------------------------------------------
template <int WP>
struct Syrk
{
  static constexpr int NB = WP / 4;
  static constexpr int NT = NB * (NB + 1) / 2;
  static constexpr int KS_ = (SAGE_CTA / NT) < 1 ? 1 : (SAGE_CTA / NT);
  static constexpr int KS = KS_ > 8 ? 8 : KS_;
  float acc[16];
  int bi, bj, ks;
  bool active;

  __device__ __forceinline__ void init()
  {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      acc[i] = 0.f;
    const int t = threadIdx.x;
    ks = t / NT;
    active = ks < KS;
    int tile = t - ks * NT;
    // tile -> (bi <= bj), row-major over the upper triangle
    int b = 0;
    int rem = tile;
    while (rem >= NB - b)
    {
      rem -= NB - b;
      ++b;
    }
    bi = b;
    bj = b + rem;
  }

  // rows [0, nrows) of Y (row stride WP floats, 16-byte aligned)
  __device__ __forceinline__ void accumulate(const float *Y, int nrows)
  {
    if (!active)
      return;
    for (int r = ks; r < nrows; r += KS)
    {
      const float4 a = *reinterpret_cast<const float4 *>(Y + r * WP + bi * 4);
      const float4 b = *reinterpret_cast<const float4 *>(Y + r * WP + bj * 4);
      acc[0] = fmaf(a.x, b.x, acc[0]); acc[1] = fmaf(a.x, b.y, acc[1]); acc[2] = fmaf(a.x, b.z, acc[2]); acc[3] = fmaf(a.x, b.w, acc[3]);
      acc[4] = fmaf(a.y, b.x, acc[4]); acc[5] = fmaf(a.y, b.y, acc[5]); acc[6] = fmaf(a.y, b.z, acc[6]); acc[7] = fmaf(a.y, b.w, acc[7]);
      acc[8] = fmaf(a.z, b.x, acc[8]); acc[9] = fmaf(a.z, b.y, acc[9]); acc[10] = fmaf(a.z, b.z, acc[10]); acc[11] = fmaf(a.z, b.w, acc[11]);
      acc[12] = fmaf(a.w, b.x, acc[12]); acc[13] = fmaf(a.w, b.y, acc[13]); acc[14] = fmaf(a.w, b.z, acc[14]); acc[15] = fmaf(a.w, b.w, acc[15]);
    }
  }

  // Sum the K-split groups through shared memory (scratch: NT*16 floats) and add the CTA's result to
  // the full WP x WP (row-major, both triangles of off-diagonal tiles mirrored) partial in global
  // memory.  dst is this CTA's private slice, so plain stores suffice.
  __device__ __forceinline__ void store(float *scratch, float *dst)
  {
    __syncthreads();
    for (int g = KS - 1; g >= 1; --g)
    {
      if (active && ks == g)
      {
        const int tile = threadIdx.x - ks * NT;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          scratch[tile * 16 + i] = acc[i];
      }
      __syncthreads();
      if (active && ks == 0)
      {
        const int tile = threadIdx.x;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          acc[i] += scratch[tile * 16 + i];
      }
      __syncthreads();
    }
    if (active && ks == 0)
    {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const int r = bi * 4 + i, c = bj * 4 + j;
          dst[r * WP + c] = acc[i * 4 + j];
          if (bi != bj)
            dst[c * WP + r] = acc[i * 4 + j];
        }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Tensor-core symmetric rank-k accumulator (per warp): H += Y^T Y for rows Y[r][0..WP) staged in shared
// memory with row stride ST.  mma.sync.m16n8k8 TF32 with the 3xTF32 split (lo*hi + hi*lo + hi*hi, fp32
// accumulate) keeps fp32-level accuracy; only tiles touching the upper triangle are computed, and because
// A = Y^T and B = Y are the same staged rows the B fragments double as A fragments.
//   A (16x8, row): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B (8x8, col): b0 (t, g) b1 (t+4, g)
//   C (16x8): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)       with g = lane >> 2, t = lane & 3
// ------------------------------------------------------------------------------------------------
// x = hi + lo with hi, lo representable in TF32 (sign, 8 exponent, 10 mantissa bits = the top 19 bits of an fp32 word).
// Truncation split: hi = x with the low 13 mantissa bits cleared, so x - hi is exact in fp32 and holds the next 13 bits; lo is
// that remainder truncated to TF32.  |x - hi - lo| <= 2^-20 |x|, the same order as the lo*lo term the 3xTF32 scheme drops.
// Three instructions per value (LOP3, FADD, LOP3).  `cvt.rna.tf32.f32` (round to nearest, error 4x smaller) is emulated on
// sm_100a by ~5 integer / predicate instructions per conversion: with it the conversions were 42 % of all instructions of
// the geometric kernel (profiles/r1c_instruction_mix.txt: FSETP, IMAD, LOP3, SEL, IADD3).  -DSAGE_TF32_RNA restores it.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
#ifdef SAGE_TF32_RNA
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rest = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
#else
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
#endif
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int WP, int NSPLIT = 1>
struct MmaSyrk
{
  static_assert(WP % 8 == 0, "row width must be a multiple of 8");
  static constexpr int NT8 = WP / 8;        // 8-wide column tiles
  static constexpr int MT = (NT8 + 1) / 2;  // 16-tall row tiles
  static constexpr int ST = (WP % 32 == 8 || WP % 32 == 24) ? WP : WP + 8; // conflict-free fragment loads
  static constexpr int count_tiles()
  {
    int n = 0;
    for (int mi = 0; mi < MT; ++mi)
      n += NT8 - 2 * mi;
    return n;
  }
  static constexpr int NTILES = count_tiles();
  static constexpr int TPW = (NTILES + NSPLIT - 1) / NSPLIT; // tiles per cooperating warp (tile t belongs to part t % NSPLIT)
  float acc[TPW][4];

  __device__ __forceinline__ void init()
  {
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        acc[i][k] = 0.f;
  }

  // rows must be a multiple of 8; every lane of the warp calls this.  PART selects this warp's share of the tiles.
  template <int PART = 0>
  __device__ __forceinline__ void accumulate(const float *Y, int rows, int lane)
  {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < rows; k0 += 8)
    {
      uint32_t bh[NT8][2], bl[NT8][2];
      const float *r0 = Y + (size_t)(k0 + t) * ST + g, *r1 = r0 + 4 * ST;
#pragma unroll
      for (int nj = 0; nj < NT8; ++nj)
      {
        split_tf32(r0[8 * nj], bh[nj][0], bl[nj][0]);
        split_tf32(r1[8 * nj], bh[nj][1], bl[nj][1]);
      }
      int ti = 0;
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
      {
        const bool full = 2 * mi + 1 < NT8;
        const uint32_t ah0 = bh[2 * mi][0], ah2 = bh[2 * mi][1], al0 = bl[2 * mi][0], al2 = bl[2 * mi][1];
        const uint32_t ah1 = full ? bh[full ? 2 * mi + 1 : 0][0] : 0u, ah3 = full ? bh[full ? 2 * mi + 1 : 0][1] : 0u;
        const uint32_t al1 = full ? bl[full ? 2 * mi + 1 : 0][0] : 0u, al3 = full ? bl[full ? 2 * mi + 1 : 0][1] : 0u;
#pragma unroll
        for (int nj = 2 * mi; nj < NT8; ++nj)
        {
          if (ti % NSPLIT == PART)
          {
            float(&d)[4] = acc[ti / NSPLIT];
            mma_tf32(d, al0, al1, al2, al3, bh[nj][0], bh[nj][1]);
            mma_tf32(d, ah0, ah1, ah2, ah3, bl[nj][0], bl[nj][1]);
            mma_tf32(d, ah0, ah1, ah2, ah3, bh[nj][0], bh[nj][1]);
          }
          ++ti;
        }
      }
    }
  }

  // add this warp's tiles into the WP x WP matrix Hs (shared memory, zero-initialised, exclusive access)
  template <int PART = 0>
  __device__ __forceinline__ void add_to(float *Hs, int lane) const
  {
    const int g = lane >> 2, t = lane & 3;
    int ti = 0;
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int nj = 2 * mi; nj < NT8; ++nj)
      {
        if (ti % NSPLIT == PART)
        {
          const float(&d)[4] = acc[ti / NSPLIT];
          const int r = 16 * mi + g, c = 8 * nj + 2 * t;
          if (r < WP)
          {
            Hs[r * WP + c] += d[0];
            Hs[r * WP + c + 1] += d[1];
          }
          if (r + 8 < WP)
          {
            Hs[(r + 8) * WP + c] += d[2];
            Hs[(r + 8) * WP + c + 1] += d[3];
          }
        }
        ++ti;
      }
  }

  // write (not add) this warp's tiles straight to a WP x WP matrix in global memory: with NSPLIT cooperating warps the
  // tile sets are disjoint and together cover the whole upper triangle, so no reduction or zero-fill is needed
  template <int PART = 0>
  __device__ __forceinline__ void store_tiles(float *dst, int lane) const
  {
    const int g = lane >> 2, t = lane & 3;
    int ti = 0;
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int nj = 2 * mi; nj < NT8; ++nj)
      {
        if (ti % NSPLIT == PART)
        {
          const float(&d)[4] = acc[ti / NSPLIT];
          const int r = 16 * mi + g, c = 8 * nj + 2 * t;
          if (r < WP)
            *reinterpret_cast<float2 *>(dst + r * WP + c) = make_float2(d[0], d[1]);
          if (r + 8 < WP)
            *reinterpret_cast<float2 *>(dst + (r + 8) * WP + c) = make_float2(d[2], d[3]);
        }
        ++ti;
      }
  }

  // Cut the accumulation chain.  The tensor core adds into the fp32 accumulator with truncation, not round-to-nearest: every
  // mma loses about a third of an ulp of the running sum, a bias of -3e-8 per instruction that does not average out (measured on
  // sm_100a: 30720 chained instructions put -1e-3 on the diagonal of J^T J against the fp64 sum).  So every few hundred
  // instructions the tiles are added, with ordinary fp32 adds, to a second accumulator that only this warp touches and the
  // tensor-core accumulators restart from zero.
  //  * flush_frag: the second accumulator is a fragment-ordered array in shared memory (TPW x 32 float4 per warp, zeroed by the
  //    caller; lane-contiguous 16-byte accesses, conflict-free); store_frags sums the warps' arrays into the CTA's partial.
  //  * flush_tiles: it is the WP x WP partial in global memory itself (first: nothing is there yet, plain stores).
  // Only the tiles of the upper triangle exist.
  __device__ __forceinline__ void flush_frag(float4 *S, int lane)
  {
    static_assert(NSPLIT == 1, "every warp holds all tiles");
#pragma unroll
    for (int ti = 0; ti < NTILES; ++ti)
    {
      float4 v = S[ti * 32 + lane];
      v.x += acc[ti][0], v.y += acc[ti][1], v.z += acc[ti][2], v.w += acc[ti][3];
      S[ti * 32 + lane] = v;
      acc[ti][0] = acc[ti][1] = acc[ti][2] = acc[ti][3] = 0.f;
    }
  }
  // all threads of the CTA, after a block barrier: dst[r][c] = sum over the nwarps fragment arrays (S: [nwarps][NTILES * 32] float4)
  static __device__ __forceinline__ void store_frags(const float4 *S, float *dst, int nwarps)
  {
    const float *Sf = reinterpret_cast<const float *>(S);
    for (int e = threadIdx.x; e < NTILES * 128; e += blockDim.x)
    {
      const int ti = e >> 7, l = (e >> 2) & 31, k = e & 3;
      int mi = 0, first = 0; // tile ti = (mi, nj): tiles are numbered row by row, row mi holds nj = 2 mi .. NT8 - 1
      while (ti >= first + NT8 - 2 * mi)
      {
        first += NT8 - 2 * mi;
        ++mi;
      }
      const int nj = 2 * mi + (ti - first);
      const int r = 16 * mi + (l >> 2) + ((k & 2) ? 8 : 0), c = 8 * nj + 2 * (l & 3) + (k & 1);
      float v = 0.f;
      for (int w = 0; w < nwarps; ++w)
        v += Sf[(size_t)w * NTILES * 128 + e];
      if (r < WP)
        dst[r * WP + c] = v;
    }
  }
  template <int PART = 0>
  __device__ __forceinline__ void flush_tiles(float *dst, int lane, bool first)
  {
    const int g = lane >> 2, t = lane & 3;
    int ti = 0;
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int nj = 2 * mi; nj < NT8; ++nj)
      {
        if (ti % NSPLIT == PART)
        {
          float(&d)[4] = acc[ti / NSPLIT];
          const int r = 16 * mi + g, c = 8 * nj + 2 * t;
          if (r < WP)
          {
            float2 *p = reinterpret_cast<float2 *>(dst + r * WP + c);
            float2 o = make_float2(d[0], d[1]);
            if (!first)
            {
              const float2 q = *p;
              o.x += q.x, o.y += q.y;
            }
            *p = o;
          }
          if (r + 8 < WP)
          {
            float2 *p = reinterpret_cast<float2 *>(dst + (r + 8) * WP + c);
            float2 o = make_float2(d[2], d[3]);
            if (!first)
            {
              const float2 q = *p;
              o.x += q.x, o.y += q.y;
            }
            *p = o;
          }
          d[0] = d[1] = d[2] = d[3] = 0.f;
        }
        ++ti;
      }
  }

  // NSPLIT == 1: sum the warps' full accumulators in shared memory (Hs: WP*WP floats, may alias the staging buffer once
  // every warp is done with it) and write the CTA's WP x WP partial (upper triangle valid) to dst.  Called by all threads.
  __device__ __forceinline__ void store_cta(float *Hs, float *dst, int warp, int lane, int nwarps)
  {
    for (int i = threadIdx.x; i < WP * WP; i += blockDim.x)
      Hs[i] = 0.f;
    __syncthreads();
    for (int w = 0; w < nwarps; ++w)
    {
      if (warp == w)
        add_to<0>(Hs, lane);
      __syncthreads();
    }
    for (int i = threadIdx.x; i < WP * WP; i += blockDim.x)
      dst[i] = Hs[i];
  }
};

// Opt-in attributes such as cudaFuncAttributeMaxDynamicSharedMemorySize live per device function PER DEVICE (context): a process
// that drives several GPUs must set them once on each.  Returns true the first time it is called for `mask` on the current device.
inline bool first_use_on_device(unsigned long long &mask)
{
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit)
    return false;
  mask |= bit;
  return true;
}

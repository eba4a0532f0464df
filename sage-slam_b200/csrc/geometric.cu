// geometric.cu -- fused geometric (depth-consistency) linearisation / error kernels (sm_100a).
//
// Replaces geometric_jac_error_calculate_kernel (cuda/geometric_factor_kernels.cpp:472-720),
// geometric_error_calculate_kernel (:127-218) and the ATen reductions after them (:931-947, :870-879).
//
// A CTA is a pair of warps; each warp takes batches of 32 samples.
//  * lane == sample: depth of the sample, warp into KF1, nearest mask lookup, the bilinear taps of KF1's
//    (depth, d/dx, d/dy, mask) float4 map, Cauchy weight and the 9 "small" columns of the Jacobian row
//    ([pose0 6 | scale0 | scale1 | rhs]; the pose1 block is exactly -pose0 and is expanded at the end).
//  * lane == channel quad: a group of C/4 lanes fetches the KF0 basis row and the four taps of KF1's
//    pixel-major basis [HW][C] as float4 (one fully used 128-byte line per tap for C = 32) and writes the
//    2C code columns of the row.
//  * The staged rows (width 16 + 2C, one per sample) of both warps are folded into J^T J | J^T r on the
//    tensor cores (mma.sync m16n8k8, 3xTF32, fp32 accumulate), the upper-triangular tiles split between
//    the two warps.  Nothing is written to HBM except one partial per CTA.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

struct GeoCam
{
  float fx, fy, cx, cy;
  int W, H;
};

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
  const int W = cam.W, H = cam.H;
  float *Yw = Y + (size_t)warp * 32 * ST;

  typename T::Syrk syrk;
  if constexpr (JAC)
    syrk.init();
  float err_acc = 0.f, inl_acc = 0.f;

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
    const float hx = hm.x, hy = hm.y, hz = hm.z;
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
    const float rx = fs.R10[0] * hx + fs.R10[1] * hy + fs.R10[2] * hz;
    const float ry = fs.R10[3] * hx + fs.R10[4] * hy + fs.R10[5] * hz;
    const float rz = fs.R10[6] * hx + fs.R10[7] * hy + fs.R10[8] * hz;
    const float px = d0 * rx + fs.t10[0], py = d0 * ry + fs.t10[1], pz = d0 * rz + fs.t10[2];
    const bool pos = pz > fs.eps;
    float ux = (px / pz) * cam.fx + cam.cx;
    float uy = (py / pz) * cam.fy + cam.cy;
    const int mx = (int)roundf(ux), my = (int)roundf(uy);
    const float wm = (live && pos && within(mx, my, W, H)) ? __ldg(&fs.dgm1[my * W + mx].w) : 0.f;
    const float valid = wm;
    if (valid == 0.f)
    {
      ux = 0.f; // keep everything finite; the row is multiplied by the zero weight below
      uy = 0.f;
    }
    const TapSet tp = make_tapset(ux, uy, W, H, 4); // pk & ~3 = pixel * 4 (float offset into the float4 map)
    float D1, gx, gy;
    {
      const float *pnw = reinterpret_cast<const float *>(fs.dgm1) + (tp.pk & ~3);
      const float *pne = pnw + ((tp.pk & 2) ? 4 : 0);
      const float *psw = pnw + ((tp.pk & 1) ? 4 * W : 0);
      const float *pse = psw + ((tp.pk & 2) ? 4 : 0);
      const float4 dgv = gather4(pnw, pse, psw, pne, tp.w);
      D1 = fs.dscale * dgv.x; // sampled (scaled) depth of KF1 and its gradient
      gx = fs.dscale * dgv.y;
      gy = fs.dscale * dgv.z;
    }
    const float diff = D1 - pz;
    const float md = wm * diff;
    const float e = valid != 0.f ? logf(1.0f + (md * md) / fs.loss_param) : 0.f; // :600
    err_acc += e;
    inl_acc += valid;

    if constexpr (JAC)
    {
      const bool on = valid != 0.f;
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
      float *row = Yw + (size_t)lane * ST;
      *reinterpret_cast<float4 *>(row + 0) = make_float4(sw * v[0], sw * v[1], sw * v[2], sw * p3);
      *reinterpret_cast<float4 *>(row + 4) = make_float4(sw * p4, sw * p5, sw * js0, sw * js1);
      *reinterpret_cast<float4 *>(row + 8) = make_float4(sw * diff, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4 *>(row + 12) = make_float4(0.f, 0.f, 0.f, 0.f);
      // ---------------------------------------------------------------- lane == channel quad: code columns
      const float k0 = sw * kc0, k1 = -(sw * fs.scale1); // :695-696
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
      __syncthreads();
    }
  }

  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  if constexpr (JAC)
  {
    float *dst = partH + slot * (WP * WP);
    if (warp == 0)
      syrk.template store_tiles<0>(dst, lane);
    else
      syrk.template store_tiles<1>(dst, lane); // disjoint tiles covering the upper triangle: no reduction needed
  }
  const float es = block_sum(err_acc, red);
  const float cs = block_sum(inl_acc, red);
  if (threadIdx.x == 0)
  {
    partE[slot * 2 + 0] = es;
    partE[slot * 2 + 1] = cs;
  }
}

// [AtA D*D | Atb D | error | inliers], D = 14 + 2C, scaled by weight / n_inliers; zero overlap -> 10*weight, zeros (:931-947)
template <int C, bool JAC>
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
      for (int s = 0; s < slices; ++s)
        v += partH[((size_t)slot * slices + s) * (WP * WP) + lo * WP + hi];
      o[e] = sr * scn * v * sc;
    }
  }
}

int geo_row_width(int C) { return 16 + 2 * C; }
int geo_samples_per_cta() { return GEO_WARPS * 32; }

template <int C>
static void launch_geo_c(bool jac, const GeoFactor *factors, int nfactors, const GeoCam &cam, int slices, float *partH, float *partE,
                         float *out, int out_stride, cudaStream_t stream)
{
  dim3 grid(slices, nfactors);
  if (jac)
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
int geo_ctas_per_sm(bool jac, int C)
{
  int n = 0;
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
               int slices, float *partH, float *partE, float *out, int out_stride, cudaStream_t stream)
{
  if (nfactors <= 0)
    return 0;
  GeoCam cam{fx, fy, cx, cy, W, H};
  switch (C)
  {
  case 32: launch_geo_c<32>(jac, factors, nfactors, cam, slices, partH, partE, out, out_stride, stream); return 0;
  case 16: launch_geo_c<16>(jac, factors, nfactors, cam, slices, partH, partE, out, out_stride, stream); return 0;
  case 8: launch_geo_c<8>(jac, factors, nfactors, cam, slices, partH, partE, out, out_stride, stream); return 0;
  default: return -1;
  }
}

} // namespace sage

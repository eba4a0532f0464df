// geometric.cu -- fused geometric (depth-consistency) linearisation / error kernels (sm_100a).
//
// Replaces geometric_jac_error_calculate_kernel (cuda/geometric_factor_kernels.cpp:472-720),
// geometric_error_calculate_kernel (:127-218) and the ATen reductions after them (:931-947, :870-879).
//
// A sub-warp group of C/4 lanes owns one sample: lane j holds 4 entries of the KF0 basis row and of
// the bilinearly sampled KF1 basis (one float4 per tap from the pixel-major [HW][C] layout); the
// depth map of KF1 with its gradient and mask comes as one float4 per tap (prep.cu).  The single
// Cauchy-weighted Jacobian row (width 14+2C, plus the residual as an extra column) is staged in shared
// memory and folded into J^T J | J^T r by the cooperative Syrk<> update; nothing is written to HBM
// except one partial per CTA.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

struct GeoCam
{
  float fx, fy, cx, cy;
  int W, H;
};

// staged row: [pose0 6 | pose1 6 | scale0 | scale1 | rhs | pad | code0 C | code1 C]
template <int C>
struct GeoTraits
{
  static constexpr int LPG = C / 4;
  static constexpr int GPW = 32 / LPG;
  static constexpr int SPS = GPW * (SAGE_CTA / 32);
  static constexpr int WP = 16 + 2 * C;
};

template <int C, bool JAC>
__global__ void __launch_bounds__(SAGE_CTA, 2)
geo_kernel(const GeoFactor *__restrict__ factors, const GeoCam cam, float *__restrict__ partH, float *__restrict__ partE)
{
  using T = GeoTraits<C>;
  constexpr int LPG = T::LPG, SPS = T::SPS, WP = T::WP;
  constexpr int STAGE = JAC ? SPS * WP : 4;
  constexpr int SCR = JAC ? Syrk<WP>::NT * 16 : 4;
  __shared__ __align__(16) float Y[STAGE > SCR ? STAGE : SCR];
  __shared__ GeoFactor fs;
  __shared__ float red[32];
  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.y);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(GeoFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int gl = lane % LPG;
  const int grp = (threadIdx.x >> 5) * T::GPW + lane / LPG;
  const int N = fs.N;
  const int W = cam.W, H = cam.H;

  Syrk<WP> syrk;
  if constexpr (JAC)
    syrk.init();
  float err_acc = 0.f, inl_acc = 0.f;

  for (int base = blockIdx.x * SPS; base < N; base += gridDim.x * SPS)
  {
    const int n = base + grp;
    const bool live = n < N;
    float hx = 0.f, hy = 0.f, hz = 0.f, dot = 0.f;
    int idx = 0;
    float4 c0 = f4zero();
    if (live)
    {
      const float4 hm = __ldg(fs.homo + n);
      hx = hm.x; hy = hm.y; hz = hm.z;
      idx = __ldg(fs.loc1d + n);
      c0 = ldg4(fs.basis0 + (size_t)idx * C + gl * 4);
      dot = c0.x * fs.code0[gl * 4 + 0] + c0.y * fs.code0[gl * 4 + 1] + c0.z * fs.code0[gl * 4 + 2] + c0.w * fs.code0[gl * 4 + 3];
    }
    dot = group_sum<LPG>(dot);

    float sw = 0.f, diff = 0.f, e = 0.f, valid = 0.f;
    float pose[6], js0 = 0.f, js1 = 0.f, kc0 = 0.f;
    float4 c1 = f4zero();
#pragma unroll
    for (int k = 0; k < 6; ++k)
      pose[k] = 0.f;

    if (live)
    {
      // dpt_0 = (bias + jac . code) * scale_0     (:515-521)
      const float d0 = (__ldg(fs.bias0 + idx) + dot) * fs.scale0;
      const float rx = fs.R10[0] * hx + fs.R10[1] * hy + fs.R10[2] * hz;
      const float ry = fs.R10[3] * hx + fs.R10[4] * hy + fs.R10[5] * hz;
      const float rz = fs.R10[6] * hx + fs.R10[7] * hy + fs.R10[8] * hz;
      const float px = d0 * rx + fs.t10[0], py = d0 * ry + fs.t10[1], pz = d0 * rz + fs.t10[2];
      const bool pos = pz > fs.eps;
      const float ux = (px / pz) * cam.fx + cam.cx;
      const float uy = (py / pz) * cam.fy + cam.cy;
      const int mx = (int)roundf(ux), my = (int)roundf(uy);
      const float wm = within(mx, my, W, H) ? __ldg(&fs.dgm1[my * W + mx].w) : 0.f;
      valid = pos ? wm : 0.f;
      if (valid != 0.f)
      {
        const Taps tb = make_taps(ux, uy, W, H);
        const int o = tb.y0 * W + tb.x0;
        const float4 *dg = fs.dgm1 + o;
        const float4 z4 = f4zero();
        const float4 dgv = tap_combine(tb, tb.bnw ? __ldg(dg) : z4, tb.bse ? __ldg(dg + W + 1) : z4, tb.bsw ? __ldg(dg + W) : z4,
                                       tb.bne ? __ldg(dg + 1) : z4);
        const float D1 = fs.dscale * dgv.x; // sampled (scaled) depth of KF1
        const float gx = fs.dscale * dgv.y, gy = fs.dscale * dgv.z;
        diff = D1 - pz;
        const float md = wm * diff;
        e = logf(1.0f + (md * md) / fs.loss_param); // :600
        if constexpr (JAC)
        {
          const float *b1 = fs.basis1 + (size_t)o * C + gl * 4;
          c1 = tap_combine(tb, tb.bnw ? ldg4(b1) : z4, tb.bse ? ldg4(b1 + (size_t)(W + 1) * C) : z4,
                           tb.bsw ? ldg4(b1 + (size_t)W * C) : z4, tb.bne ? ldg4(b1 + C) : z4);
          sw = wm * sqrtf(1.0f / (diff * diff + fs.loss_param)); // :690
          const float iz = 1.0f / pz;
          const float xz = px * iz, yz = py * iz;
          const float wx = d0 * (fs.R0[0] * hx + fs.R0[1] * hy + fs.R0[2] * hz) + fs.t0[0];
          const float wy = d0 * (fs.R0[3] * hx + fs.R0[4] * hy + fs.R0[5] * hz) + fs.t0[1];
          const float wz = d0 * (fs.R0[6] * hx + fs.R0[7] * hy + fs.R0[8] * hz) + fs.t0[2];
          // v = (R1^T)[2,:] - gx * (A R1^T)[0,:] - gy * (A R1^T)[1,:]  with A the 2x3 projection Jacobian (:607-608)
          float v[3];
#pragma unroll
          for (int k = 0; k < 3; ++k)
          {
            const float a0 = cam.fx * iz * fs.R1[k * 3 + 0] - cam.fx * xz * iz * fs.R1[k * 3 + 2];
            const float a1 = cam.fy * iz * fs.R1[k * 3 + 1] - cam.fy * yz * iz * fs.R1[k * 3 + 2];
            v[k] = fs.R1[k * 3 + 2] - (gx * a0 + gy * a1);
          }
          pose[0] = v[0]; pose[1] = v[1]; pose[2] = v[2];
          pose[3] = -v[1] * wz + v[2] * wy;
          pose[4] = v[0] * wz - v[2] * wx;
          pose[5] = -v[0] * wy + v[1] * wx;
          const float jdx = cam.fx * (rx * iz - px * rz * iz * iz);
          const float jdy = cam.fy * (ry * iz - py * rz * iz * iz);
          const float d1_jac_d0 = gx * jdx + gy * jdy;
          kc0 = (rz - d1_jac_d0) * fs.scale0;          // :685
          js0 = (rz - d1_jac_d0) * d0 / fs.scale0;     // :687
          js1 = -D1 / fs.scale1;                       // :688
        }
      }
    }
    if (gl == 0)
    {
      err_acc += e;
      inl_acc += valid;
    }

    if constexpr (JAC)
    {
      float *row = Y + (size_t)grp * WP;
      if (gl == 0)
      {
        *reinterpret_cast<float4 *>(row + 0) = make_float4(sw * pose[0], sw * pose[1], sw * pose[2], sw * pose[3]);
        *reinterpret_cast<float4 *>(row + 4) = make_float4(sw * pose[4], sw * pose[5], -(sw * pose[0]), -(sw * pose[1]));
        *reinterpret_cast<float4 *>(row + 8) = make_float4(-(sw * pose[2]), -(sw * pose[3]), -(sw * pose[4]), -(sw * pose[5]));
        *reinterpret_cast<float4 *>(row + 12) = make_float4(sw * js0, sw * js1, sw * diff, 0.f);
      }
      const float k0 = sw * kc0, k1 = -(sw * fs.scale1);
      *reinterpret_cast<float4 *>(row + 16 + gl * 4) = make_float4(k0 * c0.x, k0 * c0.y, k0 * c0.z, k0 * c0.w);
      *reinterpret_cast<float4 *>(row + 16 + C + gl * 4) = make_float4(k1 * c1.x, k1 * c1.y, k1 * c1.z, k1 * c1.w);
      __syncthreads();
      syrk.accumulate(Y, SPS);
      __syncthreads();
    }
  }

  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  if constexpr (JAC)
    syrk.store(Y, partH + slot * (WP * WP));
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
  if (threadIdx.x == 0)
  {
    o[base + 0] = s_e;
    o[base + 1] = n;
  }
  if constexpr (JAC)
  {
    const float sc = n > 0.f ? f.weight / n : 0.f;
    // reference order [pose0 6 | pose1 6 | code0 C | code1 C | scale0 | scale1] -> internal column
    auto icol = [](int c) -> int { return c < 12 ? c : (c < 12 + 2 * C ? 16 + (c - 12) : 12 + (c - 12 - 2 * C)); };
    for (int e = threadIdx.x; e < D * D + D; e += blockDim.x)
    {
      int r, c;
      if (e < D * D)
      {
        r = icol(e / D);
        c = icol(e % D);
      }
      else
      {
        r = icol(e - D * D);
        c = 14;
      }
      float v = 0.f;
      for (int s = 0; s < slices; ++s)
        v += partH[((size_t)slot * slices + s) * (WP * WP) + r * WP + c];
      o[e] = v * sc;
    }
  }
}

int geo_row_width(int C) { return 16 + 2 * C; }

template <int C>
static void launch_geo_c(bool jac, const GeoFactor *factors, int nfactors, const GeoCam &cam, int slices, float *partH, float *partE,
                         float *out, int out_stride, cudaStream_t stream)
{
  dim3 grid(slices, nfactors);
  if (jac)
  {
    geo_kernel<C, true><<<grid, SAGE_CTA, 0, stream>>>(factors, cam, partH, partE);
    geo_finalize_kernel<C, true><<<nfactors, 256, 0, stream>>>(factors, slices, partH, partE, out, out_stride);
  }
  else
  {
    geo_kernel<C, false><<<grid, SAGE_CTA, 0, stream>>>(factors, cam, partH, partE);
    geo_finalize_kernel<C, false><<<nfactors, 256, 0, stream>>>(factors, slices, partH, partE, out, out_stride);
  }
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

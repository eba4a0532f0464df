// reprojection.cu -- fused keypoint reprojection linearisation / error kernels (sm_100a).
//
// Replaces reprojection_jac_error_calculate_kernel (cuda/reprojection_factor_kernels.cpp:25-211),
// reprojection_error_calculate_kernel (:213-284), tracker_reproj_jac_error_calculate_kernel (:286-363),
// tracker_reproj_error_calculate_kernel (:365-419) and their ATen reductions (:513-535, :576-597).
// M <= a few hundred matches: the work is latency-bound, so one CTA handles one factor end to end
// (rows staged in shared memory, Syrk<> accumulation, normalisation and output in the same launch);
// a batched problem launches one CTA per factor.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

// staged row: mapping [pose0 6 | pose1 6 | scale | rhs | pad 2 | code C], tracker [pose 6 | 0 | rhs]
template <int C, bool JAC, bool TRK>
__global__ void __launch_bounds__(SAGE_CTA)
reproj_kernel(const ReprojFactor *__restrict__ factors, float *__restrict__ out, int out_stride)
{
  constexpr int WP = TRK ? 8 : 16 + C;
  constexpr int D = TRK ? 6 : 13 + C;
  constexpr int STEP = 64; // matches per step -> 128 rows
  constexpr int STAGE = JAC ? 2 * STEP * WP : 4;
  constexpr int SCR = JAC ? Syrk<WP>::NT * 16 : 4;
  __shared__ __align__(16) float Y[STAGE > SCR ? STAGE : SCR];
  __shared__ __align__(16) float Hs[JAC ? WP * WP : 4];
  __shared__ ReprojFactor fs;
  __shared__ float red[32];
  __shared__ float s_e, s_n;
  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.x);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(ReprojFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();
  const int M = fs.M;
  Syrk<WP> syrk;
  if constexpr (JAC)
    syrk.init();
  float err_acc = 0.f, inl_acc = 0.f;
  const float sq = sqrtf(fs.loss_param);

  for (int base = 0; base < M; base += STEP)
  {
    const int t = threadIdx.x;
    if (t < STEP)
    {
      const int m = base + t;
      float *row0 = Y + (size_t)(2 * t) * WP, *row1 = row0 + WP;
      if (m < M)
      {
        const float hx = fs.homo[m * 3 + 0], hy = fs.homo[m * 3 + 1], hz = fs.homo[m * 3 + 2];
        float d0;
        int idx = 0;
        if constexpr (TRK)
          d0 = fs.dpts0[m];
        else
        {
          idx = fs.loc1d[m];
          d0 = fs.bias0[idx];
          for (int i = 0; i < C; ++i)
            d0 += fs.basis0[(size_t)idx * C + i] * fs.code0[i];
          d0 *= fs.scale0; // :58-65
        }
        const float rx = fs.R10[0] * hx + fs.R10[1] * hy + fs.R10[2] * hz;
        const float ry = fs.R10[3] * hx + fs.R10[4] * hy + fs.R10[5] * hz;
        const float rz = fs.R10[6] * hx + fs.R10[7] * hy + fs.R10[8] * hz;
        const float px = d0 * rx + fs.t10[0], py = d0 * ry + fs.t10[1], pz = d0 * rz + fs.t10[2];
        const bool pos = pz > fs.eps;
        const float ux = (px / pz) * fs.fx + fs.cx, uy = (py / pz) * fs.fy + fs.cy;
        const float dx = fs.match2d[m * 2 + 0] - ux, dy = fs.match2d[m * 2 + 1] - uy;
        const float nx = fabsf(dx) / sq, ny = fabsf(dy) / sq;
        if (pos)
        {
          err_acc += 2.0f * (nx + ny - logf(1.0f + nx) - logf(1.0f + ny)); // :97-99
          inl_acc += 1.0f;
        }
        if constexpr (JAC)
        {
          const float wx_ = pos ? sqrtf(1.0f / (fs.loss_param * (1.0f + nx))) : 0.f; // :93-94
          const float wy_ = pos ? sqrtf(1.0f / (fs.loss_param * (1.0f + ny))) : 0.f;
          const float iz = pos ? 1.0f / pz : 0.f;
          const float xz = px * iz, yz = py * iz;
          if constexpr (TRK)
          {
            const float P0[6] = {fs.fx * iz, 0.f, -fs.fx * xz * iz, -fs.fx * xz * yz, fs.fx * (1.0f + xz * xz), -fs.fx * yz};
            const float P1[6] = {0.f, fs.fy * iz, -fs.fy * yz * iz, -fs.fy * (1.0f + yz * yz), fs.fy * xz * yz, fs.fy * xz};
#pragma unroll
            for (int k = 0; k < 6; ++k)
            {
              row0[k] = wx_ * P0[k];
              row1[k] = wy_ * P1[k];
            }
            row0[6] = 0.f; row1[6] = 0.f;
            row0[7] = wx_ * dx; row1[7] = wy_ * dy;
          }
          else
          {
            const float wx = d0 * (fs.R0[0] * hx + fs.R0[1] * hy + fs.R0[2] * hz) + fs.t0[0];
            const float wy = d0 * (fs.R0[3] * hx + fs.R0[4] * hy + fs.R0[5] * hz) + fs.t0[1];
            const float wz = d0 * (fs.R0[6] * hx + fs.R0[7] * hy + fs.R0[8] * hz) + fs.t0[2];
            float a0[3], a1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
              a0[k] = fs.fx * iz * fs.R1[k * 3 + 0] - fs.fx * xz * iz * fs.R1[k * 3 + 2];
              a1[k] = fs.fy * iz * fs.R1[k * 3 + 1] - fs.fy * yz * iz * fs.R1[k * 3 + 2];
            }
            const float P0[6] = {a0[0], a0[1], a0[2], -a0[1] * wz + a0[2] * wy, a0[0] * wz - a0[2] * wx, -a0[0] * wy + a0[1] * wx};
            const float P1[6] = {a1[0], a1[1], a1[2], -a1[1] * wz + a1[2] * wy, a1[0] * wz - a1[2] * wx, -a1[0] * wy + a1[1] * wx};
#pragma unroll
            for (int k = 0; k < 6; ++k)
            {
              row0[k] = wx_ * P0[k];
              row1[k] = wy_ * P1[k];
              row0[6 + k] = -(wx_ * P0[k]);
              row1[6 + k] = -(wy_ * P1[k]);
            }
            const float jdx = fs.fx * (rx * iz - px * rz * iz * iz);
            const float jdy = fs.fy * (ry * iz - py * rz * iz * iz);
            row0[12] = wx_ * (jdx * d0 / fs.scale0); // :185
            row1[12] = wy_ * (jdy * d0 / fs.scale0);
            row0[13] = wx_ * dx; row1[13] = wy_ * dy;
            row0[14] = row0[15] = row1[14] = row1[15] = 0.f;
            for (int i = 0; i < C; ++i)
            {
              const float bi = fs.basis0[(size_t)idx * C + i];
              row0[16 + i] = wx_ * (jdx * fs.scale0 * bi); // :181-182
              row1[16 + i] = wy_ * (jdy * fs.scale0 * bi);
            }
          }
        }
      }
      else if constexpr (JAC)
      {
        for (int k = 0; k < WP; ++k)
        {
          row0[k] = 0.f;
          row1[k] = 0.f;
        }
      }
    }
    if constexpr (JAC)
    {
      __syncthreads();
      syrk.accumulate(Y, 2 * STEP);
      __syncthreads();
    }
  }
  if constexpr (JAC)
    syrk.store(Y, Hs);
  const float es = block_sum(err_acc, red);
  const float ns = block_sum(inl_acc, red);
  if (threadIdx.x == 0)
  {
    s_n = ns;
    s_e = ns > 0.f ? (fs.weight / ns) * es : fs.weight * 10.0f; // :520-535
  }
  __syncthreads();
  float *o = out + (size_t)fs.out * out_stride;
  const int eb = JAC ? D * D + D : 0;
  if (threadIdx.x == 0)
  {
    o[eb] = s_e;
    o[eb + 1] = s_n;
  }
  if constexpr (JAC)
  {
    const float sc = s_n > 0.f ? fs.weight / s_n : 0.f;
    auto icol = [](int c) -> int {
      if constexpr (TRK)
        return c;
      else
        return c < 12 ? c : (c < 12 + C ? 16 + (c - 12) : 12);
    };
    constexpr int RHS = TRK ? 7 : 13;
    for (int e = threadIdx.x; e < D * D + D; e += blockDim.x)
    {
      const int r = e < D * D ? icol(e / D) : icol(e - D * D);
      const int c = e < D * D ? icol(e % D) : RHS;
      o[e] = Hs[r * WP + c] * sc;
    }
  }
}


// tracker match-geometry: 3 rows per match [pose 6 | scale | rhs], normalised by the number of matches (torch::mean,
// K/match_geometry_factor_kernels.cpp:1380, :1405-1418); one CTA, M <= 4096.
template <bool JAC>
__global__ void __launch_bounds__(SAGE_CTA)
match_geom_kernel(const MatchGeomFactor *__restrict__ factors, float *__restrict__ out, int out_stride, int D)
{
  constexpr int WP = 8, STEP = 64;
  __shared__ __align__(16) float Y[JAC ? 3 * STEP * WP : 4];
  __shared__ __align__(16) float Hs[JAC ? WP * WP : 4];
  __shared__ MatchGeomFactor fs;
  __shared__ float red[32];
  __shared__ float s_e;
  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.x);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(MatchGeomFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();
  const int M = fs.M;
  Syrk<WP> syrk;
  if constexpr (JAC)
    syrk.init();
  float err_acc = 0.f;
  const float sq = sqrtf(fs.loss_param);
  for (int base = 0; base < M; base += STEP)
  {
    const int tI = threadIdx.x;
    if (tI < STEP)
    {
      const int m = base + tI;
      float *rows = Y + (size_t)(3 * tI) * WP;
      if (m < M)
      {
        const float hx = fs.homo0[m * 3 + 0], hy = fs.homo0[m * 3 + 1], hz = fs.homo0[m * 3 + 2];
        const float d0 = fs.dmul * fs.dpts0[m], d1 = fs.dpts1[m];
        const float r[3] = {fs.R[0] * hx + fs.R[1] * hy + fs.R[2] * hz, fs.R[3] * hx + fs.R[4] * hy + fs.R[5] * hz,
                            fs.R[6] * hx + fs.R[7] * hy + fs.R[8] * hz};
        const float p[3] = {d0 * r[0] + fs.t[0], d0 * r[1] + fs.t[1], d0 * r[2] + fs.t[2]};
        const float Jp[3][6] = {{1.f, 0.f, 0.f, 0.f, p[2], -p[1]}, {0.f, 1.f, 0.f, -p[2], 0.f, p[0]}, {0.f, 0.f, 1.f, p[1], -p[0], 0.f}};
        float e = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          const float diff = d1 * fs.homo1[m * 3 + i] - p[i];
          const float nrm = fabsf(diff) / sq;
          e += nrm - logf(1.0f + nrm);
          if constexpr (JAC)
          {
            const float w = sqrtf(1.0f / (fs.loss_param * (1.0f + nrm)));
#pragma unroll
            for (int j = 0; j < 6; ++j)
              rows[i * WP + j] = w * Jp[i][j];
            rows[i * WP + 6] = fs.scale0 != 0.f ? w * (r[i] * d0 / fs.scale0) : 0.f;
            rows[i * WP + 7] = w * diff;
          }
        }
        err_acc += 2.0f * e;
      }
      else if constexpr (JAC)
      {
        for (int k = 0; k < 3 * WP; ++k)
          rows[k] = 0.f;
      }
    }
    if constexpr (JAC)
    {
      __syncthreads();
      syrk.accumulate(Y, 3 * STEP);
      __syncthreads();
    }
  }
  if constexpr (JAC)
    syrk.store(Y, Hs);
  const float es = block_sum(err_acc, red);
  if (threadIdx.x == 0)
    s_e = M > 0 ? fs.weight * es / (float)M : 0.f;
  __syncthreads();
  float *o = out + (size_t)fs.out * out_stride;
  const int eb = JAC ? D * D + D : 0;
  if (threadIdx.x == 0)
  {
    o[eb] = s_e;
    o[eb + 1] = (float)M;
  }
  if constexpr (JAC)
  {
    const float sc = M > 0 ? fs.weight / (float)M : 0.f;
    for (int e = threadIdx.x; e < D * D + D; e += blockDim.x)
    {
      const int r = e < D * D ? e / D : e - D * D;
      const int c = e < D * D ? e % D : 7;
      o[e] = Hs[r * WP + c] * sc;
    }
  }
}

int launch_match_geom(bool jac, const MatchGeomFactor *factors, int nfactors, float *out, int out_stride, int D, cudaStream_t stream)
{
  if (nfactors <= 0)
    return 0;
  if (jac)
    match_geom_kernel<true><<<nfactors, SAGE_CTA, 0, stream>>>(factors, out, out_stride, D);
  else
    match_geom_kernel<false><<<nfactors, SAGE_CTA, 0, stream>>>(factors, out, out_stride, D);
  return 0;
}

template <int C>
static void launch_reproj_c(bool jac, bool tracker, const ReprojFactor *f, int nf, float *out, int out_stride, cudaStream_t s)
{
  if (tracker)
  {
    if (jac)
      reproj_kernel<C, true, true><<<nf, SAGE_CTA, 0, s>>>(f, out, out_stride);
    else
      reproj_kernel<C, false, true><<<nf, SAGE_CTA, 0, s>>>(f, out, out_stride);
  }
  else
  {
    if (jac)
      reproj_kernel<C, true, false><<<nf, SAGE_CTA, 0, s>>>(f, out, out_stride);
    else
      reproj_kernel<C, false, false><<<nf, SAGE_CTA, 0, s>>>(f, out, out_stride);
  }
}

int launch_reproj(bool jac, bool tracker, int C, const ReprojFactor *factors, int nfactors, float *out, int out_stride,
                  cudaStream_t stream)
{
  if (nfactors <= 0)
    return 0;
  if (tracker)
  {
    launch_reproj_c<8>(jac, true, factors, nfactors, out, out_stride, stream);
    return 0;
  }
  switch (C)
  {
  case 32: launch_reproj_c<32>(jac, false, factors, nfactors, out, out_stride, stream); return 0;
  case 16: launch_reproj_c<16>(jac, false, factors, nfactors, out, out_stride, stream); return 0;
  case 8: launch_reproj_c<8>(jac, false, factors, nfactors, out, out_stride, stream); return 0;
  default: return -1;
  }
}

} // namespace sage

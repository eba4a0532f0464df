// match_geometry.cu -- mapping-side match-geometry and loop-closure match-geometry factors (sm_100a).
//
// Replaces match_geometry_jac_error_calculate_kernel_{fair,l2,huber,unbiased}
// (cuda/match_geometry_factor_kernels.cpp:577-726, :730-868, :872-1039, :421-573), the four
// match_geometry_error_calculate_kernel_* (:1090-1359), loop_mg_jac_error_calculate_kernel (:296-417),
// loop_mg_error_calculate_kernel (:1043-1086) and the ATen reductions behind them (:1479-1565, :1567-1823).
// 3 residual rows per match (x, y, z of the 3-D point difference), M <= a few hundred matches: one CTA per
// factor end to end, like reprojection.cu; the linearising form uses one thread per (match, axis) row.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

// depths of the match (after scale) and the unscaled values the scale columns need
template <int C>
__device__ __forceinline__ void mg_depths(const MapMatchGeomFactor &fs, int m, float &d0, float &d1, float &u0, float &u1)
{
  if constexpr (C == 0)
  {
    u0 = fs.dpts0[m];
    u1 = fs.dpts1[m];
    d0 = u0 * fs.scale0; // :315-316
    d1 = u1 * fs.scale1;
  }
  else
  {
    const int i0 = fs.loc0[m], i1 = fs.loc1[m];
    float a = fs.bias0[i0], b = fs.bias1[i1];
    for (int i = 0; i < C; ++i)
      a += fs.basis0[(size_t)i0 * C + i] * fs.code0[i];
    for (int i = 0; i < C; ++i)
      b += fs.basis1[(size_t)i1 * C + i] * fs.code1[i];
    u0 = a;
    u1 = b;
    if (fs.loss_type == 3)
    {
      const float sum = fs.scale0 + fs.scale1; // :447-462
      d0 = a * fs.scale0 / sum;
      d1 = b * fs.scale1 / sum;
    }
    else
    {
      d0 = a * fs.scale0; // :602-616
      d1 = b * fs.scale1;
    }
  }
}

// staged row: [pose0 6 | pose1 6 | scale0 | scale1 | rhs | pad | code0 C | code1 C]
template <int C, bool JAC>
__global__ void __launch_bounds__(SAGE_CTA)
map_match_geom_kernel(const MapMatchGeomFactor *__restrict__ factors, float *__restrict__ out, int out_stride)
{
  constexpr int WP = 16 + 2 * C;
  constexpr int D = 14 + 2 * C;
  constexpr int STEP = C > 16 ? 16 : 64; // matches per step -> 3 * STEP rows
  constexpr int STAGE = JAC ? 3 * STEP * WP : 4;
  constexpr int SCR = JAC ? Syrk<WP>::NT * 16 : 4;
  __shared__ __align__(16) float Y[STAGE > SCR ? STAGE : SCR];
  __shared__ __align__(16) float Hs[JAC ? WP * WP : 4];
  __shared__ MapMatchGeomFactor fs;
  __shared__ float red[32];
  __shared__ float s_e;
  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.x);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(MapMatchGeomFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();
  const int M = fs.M;
  const float c = fs.loss_param, sq = sqrtf(c);
  float err_acc = 0.f;

  if constexpr (!JAC)
  {
    for (int m = threadIdx.x; m < M; m += blockDim.x)
    {
      float d0, d1, u0, u1;
      mg_depths<C>(fs, m, d0, d1, u0, u1);
      const float hx = fs.homo0[m * 3 + 0], hy = fs.homo0[m * 3 + 1], hz = fs.homo0[m * 3 + 2];
      float diff[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        const float p = d0 * (fs.R10[i * 3 + 0] * hx + fs.R10[i * 3 + 1] * hy + fs.R10[i * 3 + 2] * hz) + fs.t10[i];
        diff[i] = d1 * fs.homo1[m * 3 + i] - p;
      }
      float e;
      if (fs.loss_type == 1)
        e = diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2];
      else if (fs.loss_type == 2)
      {
        // the reference expression `err + (sq <= c) ? sq : 2 sqrt(c sq) - c` (:1349-1356) is kept as it parses
        e = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          const float s2 = diff[i] * diff[i];
          const float cond = e + (s2 <= c ? 1.f : 0.f);
          e = cond != 0.f ? s2 : (2.0f * sqrtf(c * s2) - c);
        }
      }
      else
      {
        const float n0 = fabsf(diff[0]) / sq, n1 = fabsf(diff[1]) / sq, n2 = fabsf(diff[2]) / sq;
        e = 2.0f * (n0 + n1 + n2 - logf(1.0f + n0) - logf(1.0f + n1) - logf(1.0f + n2));
      }
      err_acc += e;
    }
  }
  else
  {
    Syrk<WP> syrk;
    syrk.init();
    for (int base = 0; base < M; base += STEP)
    {
      const int tI = threadIdx.x;
      if (tI < 3 * STEP)
      {
        const int m = base + tI / 3, i = tI % 3; // one thread per residual row
        float *row = Y + (size_t)tI * WP;
        if (m < M)
        {
          float d0, d1, u0, u1;
          mg_depths<C>(fs, m, d0, d1, u0, u1);
          const float hx = fs.homo0[m * 3 + 0], hy = fs.homo0[m * 3 + 1], hz = fs.homo0[m * 3 + 2];
          const float h1 = fs.homo1[m * 3 + i];
          const float r = fs.R10[i * 3 + 0] * hx + fs.R10[i * 3 + 1] * hy + fs.R10[i * 3 + 2] * hz;
          const float diff = d1 * h1 - (d0 * r + fs.t10[i]);
          float pw[3];
#pragma unroll
          for (int k = 0; k < 3; ++k)
            pw[k] = d0 * (fs.R0[k * 3 + 0] * hx + fs.R0[k * 3 + 1] * hy + fs.R0[k * 3 + 2] * hz) + fs.t0[k];
          float w, e;
          if (fs.loss_type == 1)
          {
            w = 1.f;
            e = diff * diff;
          }
          else if (fs.loss_type == 2)
          {
            const float s2 = diff * diff;
            e = s2 <= c ? s2 : (2.0f * sqrtf(c * s2) - c); // :943-957
            w = fminf(1.0f, sqrtf(c / s2));                 // :959-961
          }
          else
          {
            const float nrm = fabsf(diff) / sq;
            e = 2.0f * (nrm - logf(1.0f + nrm));
            w = sqrtf(1.0f / (c * (1.0f + nrm)));
          }
          err_acc += e;
          // d p1 / d pose1 (10.3.6) and d p1 / d pose0 = R1^T [I | -[p_w]x]  (:668-702)
          const float a0 = fs.R1[0 * 3 + i], a1 = fs.R1[1 * 3 + i], a2 = fs.R1[2 * 3 + i];
          row[0] = w * a0;
          row[1] = w * a1;
          row[2] = w * a2;
          row[3] = w * (-a1 * pw[2] + a2 * pw[1]);
          row[4] = w * (a0 * pw[2] - a2 * pw[0]);
          row[5] = w * (-a0 * pw[1] + a1 * pw[0]);
          row[6] = w * (-a0);
          row[7] = w * (-a1);
          row[8] = w * (-a2);
          row[9] = w * (a1 * pw[2] - a2 * pw[1]);
          row[10] = w * (-a0 * pw[2] + a2 * pw[0]);
          row[11] = w * (a0 * pw[1] - a1 * pw[0]);
          row[14] = w * diff;
          row[15] = 0.f;
          if constexpr (C == 0)
          {
            row[12] = w * (r * u0); // :412-413
            row[13] = w * (-h1 * u1);
          }
          else
          {
            const int i0 = fs.loc0[m], i1 = fs.loc1[m];
            if (fs.loss_type == 3)
            {
              const float sum = fs.scale0 + fs.scale1;
              for (int j = 0; j < C; ++j)
              {
                row[16 + j] = w * (r * fs.basis0[(size_t)i0 * C + j] * fs.scale0 / sum); // :560-563
                row[16 + C + j] = w * (-h1 * fs.basis1[(size_t)i1 * C + j] * fs.scale1 / sum);
              }
              row[12] = w * (r * d0 * fs.scale1 / (fs.scale0 * sum) + h1 * d1 / sum); // :566-569
              row[13] = w * (-r * d0 / sum - h1 * d1 * fs.scale0 / (fs.scale1 * sum));
            }
            else
            {
              for (int j = 0; j < C; ++j)
              {
                row[16 + j] = w * (r * fs.scale0 * fs.basis0[(size_t)i0 * C + j]); // :714-718
                row[16 + C + j] = w * (-h1 * fs.scale1 * fs.basis1[(size_t)i1 * C + j]);
              }
              row[12] = w * (r * d0 / fs.scale0); // :721-722
              row[13] = w * (-h1 * d1 / fs.scale1);
            }
          }
        }
        else
        {
          for (int k = 0; k < WP; ++k)
            row[k] = 0.f;
        }
      }
      __syncthreads();
      syrk.accumulate(Y, 3 * STEP);
      __syncthreads();
    }
    syrk.store(Y, Hs);
  }

  const float es = block_sum(err_acc, red);
  if (threadIdx.x == 0)
    s_e = M > 0 ? fs.weight * es / (float)M : 0.f; // weight * torch::mean (:1553, :1811)
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
    // output order [pose0 pose1 code0 code1 scale0 scale1]
    auto icol = [](int k) -> int { return k < 12 ? k : (k < 12 + 2 * C ? 16 + (k - 12) : 12 + (k - 12 - 2 * C)); };
    for (int e = threadIdx.x; e < D * D + D; e += blockDim.x)
    {
      const int r = e < D * D ? icol(e / D) : icol(e - D * D);
      const int cc = e < D * D ? icol(e % D) : 14;
      o[e] = Hs[r * WP + cc] * sc;
    }
  }
}

template <int C>
static void launch_mmg_c(bool jac, const MapMatchGeomFactor *f, int nf, float *out, int out_stride, cudaStream_t s)
{
  if (jac)
    map_match_geom_kernel<C, true><<<nf, SAGE_CTA, 0, s>>>(f, out, out_stride);
  else
    map_match_geom_kernel<C, false><<<nf, SAGE_CTA, 0, s>>>(f, out, out_stride);
}

// C == 0 selects the loop-closure form
int launch_map_match_geom(bool jac, int C, const MapMatchGeomFactor *factors, int nfactors, float *out, int out_stride, cudaStream_t stream)
{
  if (nfactors <= 0)
    return 0;
  switch (C)
  {
  case 0: launch_mmg_c<0>(jac, factors, nfactors, out, out_stride, stream); return 0;
  case 8: launch_mmg_c<8>(jac, factors, nfactors, out, out_stride, stream); return 0;
  case 16: launch_mmg_c<16>(jac, factors, nfactors, out, out_stride, stream); return 0;
  case 32: launch_mmg_c<32>(jac, factors, nfactors, out, out_stride, stream); return 0;
  default: return -1;
  }
}

} // namespace sage

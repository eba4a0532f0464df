// photometric.cu -- fused photometric linearisation / error kernels (sm_100a).
//
// Replaces, for every pyramid level at once and without materialising the Jacobian:
//   photometric_jac_error_calculate_kernel   cuda/photometric_factor_kernels.cpp:33-368   (PH_MAP_JAC)
//   photometric_error_calculate_kernel       :370-522                                    (PH_MAP_ERR)
//   tracker_photo_jac_error_calculate_kernel :524-695  / _with_scale_kernel :697-873     (PH_TRK_JAC)
//   tracker_photo_error_calculate_kernel     :875-988                                    (PH_TRK_ERR)
// plus the ATen reductions that follow them (:1139-1161, :1220-1242, :1301-1322, :1049-1057).
//
// Work decomposition: a sub-warp group of F/4 lanes owns one sample point; each lane owns 4 feature
// channels and fetches them as one float4 per bilinear tap from the channel-last pyramid
// [SP][3][F] (feature | d/dx | d/dy), i.e. every tap of every map is one fully used 128-byte line
// per group for F = 32.  Per sample the F x L residual rows J = g~^T P^ (g~: level-scaled sampled
// gradient, P^: 2 x D level-independent projection Jacobian) collapse to the 2x2 Gram matrix
// G = sum w_l g~ g~^T and b = sum w_l g~ r; a Cholesky factor of G turns them into two "virtual
// rows" of width D+1 that are staged in shared memory and accumulated into J^T J | J^T r by a
// cooperative register-tiled rank-k update (Syrk<>).  Each CTA writes one private partial; a
// second tiny kernel reduces partials in a fixed order (deterministic) and applies the
// inlier normalisation / zero-overlap fallback.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

template <int F, int C, int MODE>
struct PhotoTraits
{
  static constexpr bool kJac = (MODE == PH_MAP_JAC || MODE == PH_TRK_JAC);
  static constexpr bool kMap = (MODE == PH_MAP_JAC || MODE == PH_MAP_ERR);
  static constexpr int LPG = F / 4;           // lanes per sample
  static constexpr int GPW = 32 / LPG;        // samples per warp
  static constexpr int SPS = GPW * (SAGE_CTA / 32); // samples per CTA step
  // staged row: [pose0 6 | pose1 6 | scale | rhs | pad 2 | code C]   (mapping)   /   [pose 6 | scale | rhs] (tracker)
  static constexpr int WP = kMap ? 16 + C : 8;
};

template <int F, int C, int MODE>
__global__ void __launch_bounds__(SAGE_CTA, 2)
photo_kernel(const PhotoFactor *__restrict__ factors, const __grid_constant__ CamPyr cam, float *__restrict__ partH,
             float *__restrict__ partE)
{
  using T = PhotoTraits<F, C, MODE>;
  constexpr int LPG = T::LPG, SPS = T::SPS, WP = T::WP;
  constexpr int STAGE = T::kJac ? 2 * SPS * WP : 4;
  constexpr int SCR = T::kJac ? Syrk<WP>::NT * 16 : 4;
  __shared__ __align__(16) float Y[STAGE > SCR ? STAGE : SCR];
  __shared__ PhotoFactor fs;
  __shared__ float red[32];

  {
    const int *src = reinterpret_cast<const int *>(factors + blockIdx.y);
    int *dst = reinterpret_cast<int *>(&fs);
    for (int i = threadIdx.x; i < (int)(sizeof(PhotoFactor) / 4); i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int gl = lane % LPG;                              // lane inside the sample group
  const int grp = (threadIdx.x >> 5) * T::GPW + lane / LPG; // sample slot inside the CTA step
  const int N = fs.N;
  const int L = cam.L;

  Syrk<WP> syrk;
  if constexpr (T::kJac)
    syrk.init();
  float err_acc = 0.f, inl_acc = 0.f;

  for (int base = blockIdx.x * SPS; base < N; base += gridDim.x * SPS)
  {
    const int n = base + grp;
    const bool live = n < N;
    float Gxx = 0.f, Gxy = 0.f, Gyy = 0.f, bx = 0.f, by = 0.f, esum = 0.f;
    float valid = 0.f;
    float px0 = 0.f, py0 = 0.f, pz0 = 1.f, rx = 0.f, ry = 0.f, rz = 0.f, d0 = 0.f;
    float hx = 0.f, hy = 0.f, hz = 0.f;
    int idx = 0;
    float4 cb[(C / 4 + LPG - 1) / LPG]; // this lane's chunks of the KF0 depth-basis row
#pragma unroll
    for (int j = 0; j < (C / 4 + LPG - 1) / LPG; ++j)
      cb[j] = f4zero();

    float dot = 0.f;
    if (live)
    {
      const float4 hm = __ldg(fs.homo + n);
      hx = hm.x; hy = hm.y; hz = hm.z;
      if constexpr (T::kMap)
      {
        idx = __ldg(fs.loc1d + n);
#pragma unroll
        for (int j = 0; j < (C / 4 + LPG - 1) / LPG; ++j)
        {
          const int ch = gl + j * LPG;
          if (ch < C / 4)
          {
            cb[j] = ldg4(fs.basis0 + (size_t)idx * C + ch * 4);
            dot += cb[j].x * fs.code0[ch * 4 + 0] + cb[j].y * fs.code0[ch * 4 + 1] + cb[j].z * fs.code0[ch * 4 + 2] +
                   cb[j].w * fs.code0[ch * 4 + 3];
          }
        }
      }
    }
    if constexpr (T::kMap)
      dot = group_sum<LPG>(dot); // outside the divergent region: every lane of the warp takes part
    if (live)
    {
      // sampled_dpts_0 = scale_0 * (bias[idx] + jac[idx,:] . code)   (:1094-1095)
      if constexpr (T::kMap)
        d0 = fs.scale0 * (__ldg(fs.bias0 + idx) + dot);
      else
        d0 = __ldg(fs.dpts0 + n);
      rx = fs.R10[0] * hx + fs.R10[1] * hy + fs.R10[2] * hz;
      ry = fs.R10[3] * hx + fs.R10[4] * hy + fs.R10[5] * hz;
      rz = fs.R10[6] * hx + fs.R10[7] * hy + fs.R10[8] * hz;
      px0 = d0 * rx + fs.t10[0];
      py0 = d0 * ry + fs.t10[1];
      pz0 = d0 * rz + fs.t10[2];
      const bool pos = pz0 > fs.eps;
      const float ux = (px0 / pz0) * cam.ofx + cam.ocx;
      const float uy = (py0 / pz0) * cam.ofy + cam.ocy;
      // nearest lookup in the full-resolution mask (:158-166); CUDA round() == roundf()
      const int mx = (int)roundf(ux), my = (int)roundf(uy);
      const float wm = within(mx, my, cam.ow, cam.oh) ? __ldg(fs.mask1 + my * cam.ow + mx) : 0.f;
      valid = pos ? wm : 0.f;

      if (valid != 0.f)
      {
        // KF pixel at level 0: a1 re-derives it from the ray (:101-103), a2 from the integer index (:423-424)
        float kx, ky;
        if constexpr (MODE == PH_MAP_JAC)
        {
          kx = hx * cam.ofx + cam.ocx;
          ky = hy * cam.ofy + cam.ocy;
        }
        else if constexpr (MODE == PH_MAP_ERR)
        {
          const float fidx = (float)idx;
          kx = fmodf(fidx, (float)cam.ow);
          ky = floorf(fidx / (float)cam.ow);
        }
        for (int l = 0; l < L; ++l)
        {
          const int W = cam.w[l], H = cam.h[l];
          const float fxl = cam.fx[l], fyl = cam.fy[l];
          const float *fg1 = fs.fg1 + (size_t)cam.off[l] * (3 * F) + gl * 4;
          float4 f0;
          if constexpr (T::kMap)
          {
            const float sx = (kx + 0.5f) * fxl / cam.ofx - 0.5f;
            const float sy = (ky + 0.5f) * fyl / cam.ofy - 0.5f;
            const Taps ta = make_taps(sx, sy, W, H);
            const float *fg0 = fs.fg0 + (size_t)cam.off[l] * (3 * F) + gl * 4;
            const int o = (ta.y0 * W + ta.x0) * (3 * F);
            const float4 a = ta.bnw ? ldg4(fg0 + o) : f4zero();
            const float4 b = ta.bse ? ldg4(fg0 + o + (W + 1) * (3 * F)) : f4zero();
            const float4 c = ta.bsw ? ldg4(fg0 + o + W * (3 * F)) : f4zero();
            const float4 d = ta.bne ? ldg4(fg0 + o + (3 * F)) : f4zero();
            f0 = tap_combine(ta, a, b, c, d);
          }
          else
          {
            f0 = ldg4(fs.sfeat0 + ((size_t)l * N + n) * F + gl * 4);
          }
          const float qx = (ux + 0.5f) * fxl / cam.ofx - 0.5f;
          const float qy = (uy + 0.5f) * fyl / cam.ofy - 0.5f;
          const Taps tb = make_taps(qx, qy, W, H);
          const int o = (tb.y0 * W + tb.x0) * (3 * F);
          const float *pnw = fg1 + o, *pse = fg1 + o + (W + 1) * (3 * F), *psw = fg1 + o + W * (3 * F),
                      *pne = fg1 + o + (3 * F);
          const float4 f1 = tap_combine(tb, tb.bnw ? ldg4(pnw) : f4zero(), tb.bse ? ldg4(pse) : f4zero(),
                                        tb.bsw ? ldg4(psw) : f4zero(), tb.bne ? ldg4(pne) : f4zero());
          float4 df;
          if constexpr (MODE == PH_MAP_ERR)
            df = make_float4(f1.x - f0.x, f1.y - f0.y, f1.z - f0.z, f1.w - f0.w);
          else
            df = make_float4(f0.x - f1.x, f0.y - f1.y, f0.z - f1.z, f0.w - f1.w);
          const float wl = fs.w[l];
          esum += wl * (wm * (df.x * df.x) + wm * (df.y * df.y) + wm * (df.z * df.z) + wm * (df.w * df.w));
          if constexpr (T::kJac)
          {
            const float4 gx = tap_combine(tb, tb.bnw ? ldg4(pnw + F) : f4zero(), tb.bse ? ldg4(pse + F) : f4zero(),
                                          tb.bsw ? ldg4(psw + F) : f4zero(), tb.bne ? ldg4(pne + F) : f4zero());
            const float4 gy = tap_combine(tb, tb.bnw ? ldg4(pnw + 2 * F) : f4zero(), tb.bse ? ldg4(pse + 2 * F) : f4zero(),
                                          tb.bsw ? ldg4(psw + 2 * F) : f4zero(), tb.bne ? ldg4(pne + 2 * F) : f4zero());
            // g~ = within_mask * sampled gradient, scaled by the level focal lengths; r = within_mask * diff
            const float sxl = wm * fxl, syl = wm * fyl;
            const float ax = gx.x * sxl, ay = gy.x * syl, bxv = gx.y * sxl, byv = gy.y * syl;
            const float cx = gx.z * sxl, cy = gy.z * syl, dx = gx.w * sxl, dy = gy.w * syl;
            const float r0 = wm * df.x, r1 = wm * df.y, r2 = wm * df.z, r3 = wm * df.w;
            Gxx += wl * (ax * ax + bxv * bxv + cx * cx + dx * dx);
            Gxy += wl * (ax * ay + bxv * byv + cx * cy + dx * dy);
            Gyy += wl * (ay * ay + byv * byv + cy * cy + dy * dy);
            bx += wl * (ax * r0 + bxv * r1 + cx * r2 + dx * r3);
            by += wl * (ay * r0 + byv * r1 + cy * r2 + dy * r3);
          }
        }
      }
    }

    esum = group_sum<LPG>(esum);
    if (gl == 0)
    {
      err_acc += esum;
      inl_acc += valid;
    }

    if constexpr (T::kJac)
    {
      Gxx = group_sum<LPG>(Gxx);
      Gxy = group_sum<LPG>(Gxy);
      Gyy = group_sum<LPG>(Gyy);
      bx = group_sum<LPG>(bx);
      by = group_sum<LPG>(by);

      // Cholesky G = L L^T and L rho = b  -> virtual rows y1 = l11 P0 + l21 P1, y2 = l22 P1
      const float l11 = sqrtf(Gxx);
      const float il11 = l11 > 0.f ? 1.0f / l11 : 0.f;
      const float l21 = Gxy * il11;
      const float l22 = sqrtf(fmaxf(Gyy - l21 * l21, 0.f));
      const float il22 = l22 > 0.f ? 1.0f / l22 : 0.f;
      const float rho1 = bx * il11;
      const float rho2 = (by - l21 * rho1) * il22;

      // level-independent projection Jacobian P^ = A^ dp1/dx with A^ = [[1/z,0,-x/z^2],[0,1/z,-y/z^2]]
      const bool on = valid != 0.f;
      const float iz = on ? 1.0f / pz0 : 0.f; // invalid samples contribute exact zeros (never 0 * inf)
      const float xz = px0 * iz, yz = py0 * iz;
      float P0[14], P1[14]; // [pose0 6 | pose1 6 | scale | rhs] (mapping) ; [pose 6 | scale | rhs] uses first 8
      if constexpr (T::kMap)
      {
        // p_w = d0 R0 x~ + t0 ; dp1/dd0 = R1^T [I | -[p_w]x],  dp1/dd1 = [-R1^T | R1^T [p_w]x]   (:247-297)
        const float wx = d0 * (fs.R0[0] * hx + fs.R0[1] * hy + fs.R0[2] * hz) + fs.t0[0];
        const float wy = d0 * (fs.R0[3] * hx + fs.R0[4] * hy + fs.R0[5] * hz) + fs.t0[1];
        const float wz = d0 * (fs.R0[6] * hx + fs.R0[7] * hy + fs.R0[8] * hz) + fs.t0[2];
        // rows of A^ R1^T: a0 = (R1^T row0)/z - xz/z (R1^T row2) ...  with R1^T[i][k] = R1[k][i]
        float a0[3], a1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
          a0[k] = iz * fs.R1[k * 3 + 0] - xz * iz * fs.R1[k * 3 + 2];
          a1[k] = iz * fs.R1[k * 3 + 1] - yz * iz * fs.R1[k * 3 + 2];
        }
        // pose0 columns: [a | a x-product with p_w]: (A^ R1^T)[I | -[pw]x]; -[pw]x = [[0,wz,-wy],[-wz,0,wx],[wy,-wx,0]]
        P0[0] = a0[0]; P0[1] = a0[1]; P0[2] = a0[2];
        P0[3] = -a0[1] * wz + a0[2] * wy;
        P0[4] = a0[0] * wz - a0[2] * wx;
        P0[5] = -a0[0] * wy + a0[1] * wx;
        P1[0] = a1[0]; P1[1] = a1[1]; P1[2] = a1[2];
        P1[3] = -a1[1] * wz + a1[2] * wy;
        P1[4] = a1[0] * wz - a1[2] * wx;
        P1[5] = -a1[0] * wy + a1[1] * wx;
#pragma unroll
        for (int k = 0; k < 6; ++k)
        {
          P0[6 + k] = -P0[k];
          P1[6 + k] = -P1[k];
        }
      }
      else
      {
        // closed-form Jacobian w.r.t. the left-perturbed relative pose (:680-681)
        P0[0] = iz; P0[1] = 0.f; P0[2] = -xz * iz; P0[3] = -xz * yz; P0[4] = 1.0f + xz * xz; P0[5] = -yz;
        P1[0] = 0.f; P1[1] = iz; P1[2] = -yz * iz; P1[3] = -(1.0f + yz * yz); P1[4] = xz * yz; P1[5] = xz;
      }
      // d pi / d depth (:324-325) without the focal length, and the scale column (:335)
      const float jdx = rx * iz - px0 * rz * iz * iz;
      const float jdy = ry * iz - py0 * rz * iz * iz;
      constexpr int SC = T::kMap ? 12 : 6;
      P0[SC] = jdx * d0 / fs.scale0;
      P1[SC] = jdy * d0 / fs.scale0;
      if constexpr (MODE == PH_TRK_JAC)
      {
        if (fs.scale0 == 0.f) // 6-DoF tracker form: no scale column
        {
          P0[SC] = 0.f;
          P1[SC] = 0.f;
        }
      }

      float *row = Y + (size_t)(2 * grp) * WP;
      {
        // lane 0 -> virtual row 1, lane 1 -> virtual row 2 (same arithmetic, different coefficients)
        const float ca = on ? (gl == 0 ? l11 : 0.f) : 0.f;
        const float cbv = on ? (gl == 0 ? l21 : l22) : 0.f;
        const float rr = on ? (gl == 0 ? rho1 : rho2) : 0.f;
        if (gl < 2)
        {
          float *dst = row + gl * WP;
          if constexpr (T::kMap)
          {
            float v[16];
#pragma unroll
            for (int k = 0; k < 13; ++k)
              v[k] = ca * P0[k] + cbv * P1[k];
            v[13] = rr; v[14] = 0.f; v[15] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          else
          {
            float v[8];
#pragma unroll
            for (int k = 0; k < 7; ++k)
              v[k] = ca * P0[k] + cbv * P1[k];
            v[7] = rr;
            *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      if constexpr (T::kMap)
      {
        // code columns: P[:, code_i] = jd * scale0 * basis_i  (:331-332)  -> y1 = k1 c, y2 = k2 c
        const float k1 = on ? (l11 * jdx + l21 * jdy) * fs.scale0 : 0.f;
        const float k2 = on ? (l22 * jdy) * fs.scale0 : 0.f;
#pragma unroll
        for (int j = 0; j < (C / 4 + LPG - 1) / LPG; ++j)
        {
          const int ch = gl + j * LPG;
          if (ch < C / 4)
          {
            *reinterpret_cast<float4 *>(row + 16 + ch * 4) = make_float4(k1 * cb[j].x, k1 * cb[j].y, k1 * cb[j].z, k1 * cb[j].w);
            *reinterpret_cast<float4 *>(row + WP + 16 + ch * 4) = make_float4(k2 * cb[j].x, k2 * cb[j].y, k2 * cb[j].z, k2 * cb[j].w);
          }
        }
      }
      __syncthreads();
      syrk.accumulate(Y, 2 * SPS);
      __syncthreads();
    }
  }

  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x; // partials are indexed by launch-local factor
  if constexpr (T::kJac)
    syrk.store(Y, partH + slot * (WP * WP));
  const float e = block_sum(err_acc, red);
  const float c = block_sum(inl_acc, red);
  if (threadIdx.x == 0)
  {
    partE[slot * 2 + 0] = e;
    partE[slot * 2 + 1] = c;
  }
}

// Reduce the per-CTA partials of one factor in a fixed order and emit the reference outputs:
//   [AtA D*D | Atb D | error | inliers]  with AtA = (1/n) sum, error = sum/n, or the zero-overlap fallback
//   (error = 10 * sum w_l, zeros)   photometric_factor_kernels.cpp:1139-1161
template <int C, int MODE>
__global__ void photo_finalize_kernel(const PhotoFactor *__restrict__ factors, int nlevels, int slices, const float *__restrict__ partH,
                                      const float *__restrict__ partE, float *__restrict__ out, int out_stride, int D)
{
  constexpr bool kJac = (MODE == PH_MAP_JAC || MODE == PH_TRK_JAC);
  constexpr bool kMap = (MODE == PH_MAP_JAC || MODE == PH_MAP_ERR);
  constexpr int WP = kMap ? 16 + C : 8;
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
    // internal column of reference column c
    auto icol = [&](int c) -> int {
      if constexpr (kMap)
        return c < 12 ? c : (c < 12 + C ? 16 + (c - 12) : 12);
      else
        return c;
    };
    constexpr int RHS = kMap ? 13 : 7;
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
        c = RHS;
      }
      float v = 0.f;
      for (int s = 0; s < slices; ++s)
        v += partH[((size_t)slot * slices + s) * (WP * WP) + r * WP + c];
      o[e] = v * inv;
    }
  }
}

template <int F, int C, int MODE>
static void launch_photo_t(const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH, float *partE,
                           float *out, int out_stride, int D, cudaStream_t stream)
{
  dim3 grid(slices, nfactors);
  photo_kernel<F, C, MODE><<<grid, SAGE_CTA, 0, stream>>>(factors, cam, partH, partE);
  photo_finalize_kernel<C, MODE><<<nfactors, 256, 0, stream>>>(factors, cam.L, slices, partH, partE, out, out_stride, D);
}

template <int F, int C>
static void launch_photo_fc(int mode, const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH,
                            float *partE, float *out, int out_stride, int D, cudaStream_t stream)
{
  switch (mode)
  {
  case PH_MAP_JAC: launch_photo_t<F, C, PH_MAP_JAC>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream); break;
  case PH_MAP_ERR: launch_photo_t<F, C, PH_MAP_ERR>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream); break;
  case PH_TRK_JAC: launch_photo_t<F, C, PH_TRK_JAC>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream); break;
  default: launch_photo_t<F, C, PH_TRK_ERR>(factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream); break;
  }
}

int photo_row_width(int mode, int C) { return (mode == PH_MAP_JAC || mode == PH_MAP_ERR) ? 16 + C : 8; }

// returns 0 on success, -1 for an unsupported (F, C)
int launch_photo(int mode, int F, int C, const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH,
                 float *partE, float *out, int out_stride, int D, cudaStream_t stream)
{
  if (nfactors <= 0)
    return 0;
#define SAGE_CASE(FF, CC)                                                                                       \
  if (F == FF && C == CC)                                                                                       \
  {                                                                                                             \
    launch_photo_fc<FF, CC>(mode, factors, nfactors, cam, slices, partH, partE, out, out_stride, D, stream);   \
    return 0;                                                                                                   \
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

// prep.cu -- once-per-keyframe re-layout and per-state depth-map preparation (sm_100a).
//
// The reference hands the factor kernels channel-major maps: feat_map_pyramid [F][SP],
// feat_map_grad_pyramid [2][F][SP] (core/mapping/mapper.cpp:1421-1424) and the depth basis as a
// (1, HW)-strided view of the net's [C][H][W] output (core/network/code_depth_network.cpp:38-39), so a
// pixel's channels sit in F (or C) different cache lines.  These kernels rewrite them ONCE per keyframe
// into channel-last form so that every bilinear tap is one contiguous, fully used line:
//     fg    [SP][3][F]   (feature | d/dx | d/dy)
//     basis [HW][C]
// and build, per state update, the depth map of a keyframe with its central-difference gradient
// (what GeometricFactor::ComputeJacobianAndError re-does on every call,
// core/gtsam/geometric_factor.cpp:317-320) packed as float4 (D, dD/dx, dD/dy, mask).
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

// generic tiled transpose: out[r][c] = in[r * sr + c * sc], r < R, c < Cn
__global__ void transpose_strided_kernel(const float *__restrict__ in, long sr, long sc, float *__restrict__ out, long R, int Cn)
{
  __shared__ float tile[32][33];
  const long r0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const long r = r0 + threadIdx.x;
    const int c = c0 + j;
    tile[j][threadIdx.x] = (r < R && c < Cn) ? in[r * sr + (long)c * sc] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const long r = r0 + j;
    const int c = c0 + threadIdx.x;
    if (r < R && c < Cn)
      out[r * Cn + c] = tile[threadIdx.x][j];
  }
}

// out[p * pitch + coff + c] = in[c * SP + p]: channel-major [F][SP] -> channel-last with an output pitch
__global__ void relayout_map_kernel(const float *__restrict__ in, float *__restrict__ out, long SP, int F, int pitch, int coff)
{
  __shared__ float tile[32][33];
  const long p0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const long p = p0 + threadIdx.x;
    const int c = c0 + j;
    tile[j][threadIdx.x] = (p < SP && c < F) ? in[(long)c * SP + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const long p = p0 + j;
    const int c = c0 + threadIdx.x;
    if (p < SP && c < F)
      out[p * pitch + coff + c] = tile[threadIdx.x][j];
  }
}

// fg[p][m][c] = (m == 0 ? feat[c][p] : grad[m-1][c][p]); grad may be null (gradient planes left untouched)
void launch_relayout_fg(const float *feat, const float *grad, float *fg, int F, long SP, cudaStream_t stream)
{
  dim3 block(32, 8);
  dim3 grid((unsigned)((SP + 31) / 32), (F + 31) / 32);
  relayout_map_kernel<<<grid, block, 0, stream>>>(feat, fg, SP, F, 3 * F, 0);
  if (grad)
  {
    relayout_map_kernel<<<grid, block, 0, stream>>>(grad, fg, SP, F, 3 * F, F);
    relayout_map_kernel<<<grid, block, 0, stream>>>(grad + (long)F * SP, fg, SP, F, 3 * F, 2 * F);
  }
}

void launch_relayout_basis(const float *jac, long stride_row, long stride_col, float *basis, int HW, int C, cudaStream_t stream)
{
  dim3 block(32, 8);
  dim3 grid((HW + 31) / 32, (C + 31) / 32);
  transpose_strided_kernel<<<grid, block, 0, stream>>>(jac, stride_row, stride_col, basis, HW, C);
}

// int64 -> int32; an index outside [0, HW) can only come from a corrupt caller array: it is clamped (never dereferenced out of
// bounds) and counted, the keyframe constructor turns a non-zero count into an error
__global__ void convert_loc_kernel(const int64_t *__restrict__ a, int *__restrict__ b, int N, int HW, int *__restrict__ bad)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
  {
    const int64_t v = a[i];
    if (v < 0 || v >= HW)
      atomicAdd(bad, 1);
    b[i] = (int)(v < 0 ? 0 : (v >= HW ? HW - 1 : v));
  }
}
void launch_convert_loc(const int64_t *loc64, int *loc32, int N, int HW, int *bad, cudaStream_t stream)
{
  if (N > 0)
    convert_loc_kernel<<<(N + 255) / 256, 256, 0, stream>>>(loc64, loc32, N, HW, bad);
}

__global__ void pack_homo_kernel(const float *__restrict__ h3, float4 *__restrict__ h4, int N)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
    h4[i] = make_float4(h3[i * 3 + 0], h3[i * 3 + 1], h3[i * 3 + 2], 0.f);
}
void launch_pack_homo(const float *homo3, float4 *homo4, int N, cudaStream_t stream)
{
  if (N > 0)
    pack_homo_kernel<<<(N + 255) / 256, 256, 0, stream>>>(homo3, homo4, N);
}

// sorted[i] = src[perm[i]] for the sample arrays (tile-major order of the staged photometric kernels)
__global__ void permute_samples_kernel(const int *__restrict__ perm, const int *__restrict__ loc, const float4 *__restrict__ homo,
                                       int *__restrict__ loc_s, float4 *__restrict__ homo_s, int N)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
  {
    const int j = perm[i];
    loc_s[i] = loc[j];
    homo_s[i] = homo[j];
  }
}
void launch_permute_samples(const int *perm, const int *loc, const float4 *homo, int *loc_s, float4 *homo_s, int N, cudaStream_t stream)
{
  if (N > 0)
    permute_samples_kernel<<<(N + 255) / 256, 256, 0, stream>>>(perm, loc, homo, loc_s, homo_s, N);
}

// D[p] = bias[p] + basis[p,:] . code      (UpdateDepth without the scale, mapping_utils.h:216-222)
__global__ void depth_unscaled_kernel(const float *__restrict__ bias, const float *__restrict__ basis, const float *__restrict__ code,
                                      float *__restrict__ D, int HW, int C)
{
  __shared__ float sc[SAGE_MAX_CODE];
  if (threadIdx.x < C)
    sc[threadIdx.x] = code[threadIdx.x];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW)
    return;
  float acc = 0.f;
  const float4 *row = reinterpret_cast<const float4 *>(basis + (size_t)p * C);
  for (int j = 0; j < C / 4; ++j)
  {
    const float4 v = __ldg(row + j);
    acc += v.x * sc[4 * j] + v.y * sc[4 * j + 1] + v.z * sc[4 * j + 2] + v.w * sc[4 * j + 3];
  }
  D[p] = __ldg(bias + p) + acc;
}

// central differences with replicate padding (ComputeSpatialGrad, mapping_utils.h:236-252)
__global__ void depth_pack_kernel(const float *__restrict__ D, const float *__restrict__ mask, float4 *__restrict__ dgm, int H, int W)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W)
    return;
  const int y = p / W, x = p - y * W;
  const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
  const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
  const float gx = 0.5f * (D[y * W + xp] - D[y * W + xm]);
  const float gy = 0.5f * (D[yp * W + x] - D[ym * W + x]);
  dgm[p] = make_float4(D[p], gx, gy, mask[p]);
}

void launch_depth_maps(const float *bias, const float *basis, const float *code_dev, const float *mask, float4 *dgm, float *scratch,
                       int H, int W, int C, cudaStream_t stream)
{
  const int HW = H * W;
  depth_unscaled_kernel<<<(HW + 255) / 256, 256, 0, stream>>>(bias, basis, code_dev, scratch, HW, C);
  depth_pack_kernel<<<(HW + 255) / 256, 256, 0, stream>>>(scratch, mask, dgm, H, W);
}

// The tracker's one-time pre-sampling (camera_tracker.cpp:1086-1123): depth and homogeneous ray of every
// sample point and its keyframe features at every level (bilinear, zero padding, pixel-centre aligned).
template <int F>
__global__ void presample_kernel(const float *__restrict__ fg0, const float *__restrict__ bias0, const float *__restrict__ basis0,
                                 const int *__restrict__ loc1d, const float4 *__restrict__ homo, const float *__restrict__ code,
                                 float scale0, const __grid_constant__ CamPyr cam, int C, int N, float *__restrict__ out_dpts,
                                 float *__restrict__ out_homo, float *__restrict__ out_feats)
{
  constexpr int LPG = F / 4;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = t / LPG, gl = t % LPG;
  if (n >= N)
    return;
  const int idx = loc1d[n];
  if (gl == 0 && out_dpts != nullptr)
  {
    float acc = 0.f;
    for (int j = 0; j < C; ++j)
      acc += basis0[(size_t)idx * C + j] * code[j];
    out_dpts[n] = scale0 * (bias0[idx] + acc);
    const float4 h = homo[n];
    out_homo[n * 3 + 0] = h.x;
    out_homo[n * 3 + 1] = h.y;
    out_homo[n * 3 + 2] = h.z;
  }
  const float x = (float)(idx % cam.ow), y = (float)(idx / cam.ow);
  for (int l = 0; l < cam.L; ++l)
  {
    const int W = cam.w[l], H = cam.h[l];
    // grid_sample(align_corners=false): ((2(x+0.5)/W0 - 1 + 1) * W_l - 1) / 2
    const float nx = (x + 0.5f) * (2.0f / (float)cam.ow) - 1.0f, ny = (y + 0.5f) * (2.0f / (float)cam.oh) - 1.0f;
    const float sx = ((nx + 1.0f) * (float)W - 1.0f) * 0.5f, sy = ((ny + 1.0f) * (float)H - 1.0f) * 0.5f;
    const Taps ta = make_taps(sx, sy, W, H);
    const float *p = fg0 + (size_t)cam.off[l] * (3 * F) + gl * 4;
    const int o = (ta.y0 * W + ta.x0) * (3 * F);
    const float4 a = ta.bnw ? ldg4(p + o) : f4zero();
    const float4 b = ta.bse ? ldg4(p + o + (W + 1) * (3 * F)) : f4zero();
    const float4 c = ta.bsw ? ldg4(p + o + W * (3 * F)) : f4zero();
    const float4 d = ta.bne ? ldg4(p + o + (3 * F)) : f4zero();
    *reinterpret_cast<float4 *>(out_feats + ((size_t)l * N + n) * F + gl * 4) = tap_combine(ta, a, b, c, d);
  }
}

void launch_presample(const float *fg0, const float *bias0, const float *basis0, const int *loc1d, const float4 *homo,
                      const float *code_dev, float scale0, const CamPyr &cam, int F, int C, int N, float *out_dpts, float *out_homo,
                      float *out_feats, cudaStream_t stream)
{
  if (N <= 0)
    return;
  if (F == 32)
    presample_kernel<32><<<(N * 8 + 255) / 256, 256, 0, stream>>>(fg0, bias0, basis0, loc1d, homo, code_dev, scale0, cam, C, N,
                                                                  out_dpts, out_homo, out_feats);
  else
    presample_kernel<16><<<(N * 4 + 255) / 256, 256, 0, stream>>>(fg0, bias0, basis0, loc1d, homo, code_dev, scale0, cam, C, N,
                                                                  out_dpts, out_homo, out_feats);
}

} // namespace sage

// ------------------------------------------------------------------------------------------------
// Per-keyframe input builder: what Mapper::BuildFrame does with torch ops
// (core/mapping/mapper.cpp:1385-1426 GenerateGaussianPyramidWithGrad, mapping_utils.h:236-252 ComputeSpatialGrad,
// mapping_utils.cpp:321-342 GenerateMaskPyramid), fused with the channel-last re-layout: the feature net's [F,H,W]
// output goes straight into fg [SP][3][F] (feature | d/dx | d/dy) without the intermediate [F,SP] / [2,F,SP] tensors.
// ------------------------------------------------------------------------------------------------
namespace sage
{

// nearest-neighbour halving of the mask: F.interpolate(mode=nearest) picks src = floor(dst * in/out)
__global__ void mask_down_kernel(const float *__restrict__ in, float *__restrict__ out, int Hi, int Wi, int Ho, int Wo)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Wo)
    return;
  const int y = p / Wo, x = p - y * Wo;
  const int sy = min((int)floorf((float)y * ((float)Hi / (float)Ho)), Hi - 1);
  const int sx = min((int)floorf((float)x * ((float)Wi / (float)Wo)), Wi - 1);
  out[p] = in[sy * Wi + sx];
}

// level 0: fg[p][0][c] = feat[c][p]  (tiled transpose), F <= 32
__global__ void pyr_level0_kernel(const float *__restrict__ feat, float *__restrict__ fg, int HW, int F)
{
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const int p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && j < F) ? feat[(size_t)j * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
  {
    const int p = p0 + j, c = threadIdx.x;
    if (p < HW && c < F)
      fg[(size_t)p * 3 * F + c] = tile[c][j];
  }
}

// level l from level l-1: 3x3 [1 2 1]^2/16, stride 2, zero padding 1, of (feat * mask) normalised by the same
// convolution of the mask (+1e-8).  One thread per (output pixel, channel).
__global__ void pyr_down_kernel(const float *__restrict__ fg_in, const float *__restrict__ mask_in, float *__restrict__ fg_out, int Hi, int Wi,
                                int Ho, int Wo, int F)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = t % F, p = t / F;
  if (p >= Ho * Wo)
    return;
  const int y = p / Wo, x = p - y * Wo;
  float acc = 0.f, macc = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
      const int yy = 2 * y - 1 + a, xx = 2 * x - 1 + b;
      if (yy >= 0 && yy < Hi && xx >= 0 && xx < Wi)
      {
        const float k = (a == 1 ? 2.f : 1.f) * (b == 1 ? 2.f : 1.f) * 0.0625f;
        const float m = mask_in[yy * Wi + xx];
        acc += k * (fg_in[(size_t)(yy * Wi + xx) * 3 * F + c] * m);
        macc += k * m;
      }
    }
  fg_out[(size_t)p * 3 * F + c] = acc / (macc + 1.0e-8f);
}

// central differences with replicate padding of the feature plane -> d/dx, d/dy planes of the same level
__global__ void pyr_grad_kernel(float *__restrict__ fg, int H, int W, int F)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = t % F, p = t / F;
  if (p >= H * W)
    return;
  const int y = p / W, x = p - y * W;
  const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
  const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
  const size_t s = (size_t)3 * F;
  fg[(size_t)p * s + F + c] = 0.5f * (fg[(size_t)(y * W + xp) * s + c] - fg[(size_t)(y * W + xm) * s + c]);
  fg[(size_t)p * s + 2 * F + c] = 0.5f * (fg[(size_t)(yp * W + x) * s + c] - fg[(size_t)(ym * W + x) * s + c]);
}

// feat [F][H*W] (device), mask0 [H*W] (device) -> fg [SP][3][F]; mask_scratch: >= H*W floats. Returns launches.
int launch_build_pyramid(const float *feat, const float *mask0, float *fg, float *mask_scratch, const CamPyr &cam, int F, cudaStream_t stream)
{
  int launches = 0;
  const int HW0 = cam.w[0] * cam.h[0];
  pyr_level0_kernel<<<(HW0 + 31) / 32, dim3(32, 8), 0, stream>>>(feat, fg, HW0, F);
  pyr_grad_kernel<<<(HW0 * F + 255) / 256, 256, 0, stream>>>(fg, cam.h[0], cam.w[0], F);
  launches += 2;
  const float *mprev = mask0;
  // the mask of level l-1 lives in mask_scratch after the first step: ping-pong between its two halves
  float *mbuf[2] = {mask_scratch, mask_scratch + HW0 / 2 + 64};
  for (int l = 1; l < cam.L; ++l)
  {
    const int Hi = cam.h[l - 1], Wi = cam.w[l - 1], Ho = cam.h[l], Wo = cam.w[l];
    float *fin = fg + (size_t)cam.off[l - 1] * 3 * F, *fout = fg + (size_t)cam.off[l] * 3 * F;
    pyr_down_kernel<<<(Ho * Wo * F + 255) / 256, 256, 0, stream>>>(fin, mprev, fout, Hi, Wi, Ho, Wo, F);
    pyr_grad_kernel<<<(Ho * Wo * F + 255) / 256, 256, 0, stream>>>(fout, Ho, Wo, F);
    launches += 2;
    if (l + 1 < cam.L)
    {
      float *mnext = mbuf[l & 1];
      mask_down_kernel<<<(Ho * Wo + 255) / 256, 256, 0, stream>>>(mprev, mnext, Hi, Wi, Ho, Wo);
      ++launches;
      mprev = mnext;
    }
  }
  return launches;
}

} // namespace sage

// api.cu -- C ABI (include/sage_ba.h): context, keyframes, the single-factor operator entry points that
// replace the reference's df::*_calculate functions, and the tracker's LM loop.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <algorithm>
#include <mutex>

#include "sage_internal.h"

using namespace sage;



namespace sage
{

CamPyr make_campyr(const sage_ba_camera &cam, int levels, sage_ba_camera *cams_out)
{
  // CameraPyramid<float> (common/camera_pyramid.h:18-32) + ResizeViewport (pinhole_camera_impl.h:120-132)
  CamPyr p;
  memset(&p, 0, sizeof(p));
  p.L = levels;
  sage_ba_camera c = cam;
  int off = 0;
  for (int i = 0; i < levels; ++i)
  {
    if (i != 0)
    {
      const float nw = (float)((size_t)c.width / 2), nh = (float)((size_t)c.height / 2);
      const float xr = nw / c.width, yr = nh / c.height;
      c.fx *= xr;
      c.fy *= yr;
      c.u0 *= xr;
      c.v0 *= yr;
      c.width = nw;
      c.height = nh;
    }
    if (cams_out)
      cams_out[i] = c;
    p.fx[i] = c.fx;
    p.fy[i] = c.fy;
    p.w[i] = (int)c.width;
    p.h[i] = (int)c.height;
    p.off[i] = off;
    off += p.w[i] * p.h[i];
  }
  p.ofx = cam.fx;
  p.ofy = cam.fy;
  p.ocx = cam.u0;
  p.ocy = cam.v0;
  p.ow = (int)cam.width;
  p.oh = (int)cam.height;
  return p;
}

static int pick_slices(const sage_ba_context *ctx, int N, int samples_per_step)
{
  const int steps = (N + samples_per_step - 1) / samples_per_step;
  if (const char *e = getenv("SAGE_BA_SLICES")) // diagnostics: accumulation-length studies
    return std::max(1, std::min(steps, atoi(e)));
  return std::max(1, std::min(steps, 3 * ctx->num_sms));
}

// run one photometric-type factor and bring [AtA | Atb | error | inliers] to the host
static void run_photo_single(sage_ba_context *ctx, int mode, int F, int C, const PhotoFactor &f, const CamPyr &pyr, int D, float *AtA,
                             float *Atb, float *error, float *n_inl)
{
  const bool jac = (mode == PH_MAP_JAC || mode == PH_TRK_JAC);
  const int WP = photo_row_width(mode, C);
  // every warp of a CTA should see a few 32-sample batches so the end-of-CTA reduction is amortised
  const int slices = pick_slices(ctx, f.N, 4 * photo_samples_per_cta());
  const size_t nout = (jac ? (size_t)D * D + D : 0) + 2;
  float *partH = ctx->partH.ensure((size_t)slices * photo_partial_floats(mode, C));
  float *partE = ctx->partE.ensure((size_t)slices * 2);
  float *out = ctx->out.ensure(nout);
  PhotoFactor *df = reinterpret_cast<PhotoFactor *>(ctx->factor.ensure(sizeof(PhotoFactor) > 1024 ? sizeof(PhotoFactor) : 1024));
  PhotoFactor *hf = reinterpret_cast<PhotoFactor *>(ctx->hfactor.ensure(1024));
  *hf = f;
  hf->out = 0;
  SAGE_CUDA(cudaMemcpyAsync(df, hf, sizeof(PhotoFactor), cudaMemcpyHostToDevice, ctx->stream));
  SAGE_CHECK(launch_photo(mode, F, C, df, 1, pyr, slices, partH, partE, out, 1, D, ctx->stream) == 0,
             "unsupported (feat_channels, code_size) combination");
  ctx->launches += 2;
  SAGE_CUDA(cudaGetLastError());
  float *h = ctx->hout.ensure(nout);
  SAGE_CUDA(cudaMemcpyAsync(h, out, nout * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SAGE_CUDA(cudaStreamSynchronize(ctx->stream));
  if (jac)
  {
    memcpy(AtA, h, sizeof(float) * D * D);
    memcpy(Atb, h + D * D, sizeof(float) * D);
  }
  *error = h[nout - 2];
  if (n_inl)
    *n_inl = h[nout - 1];
}

static void fill_pose(float *dR, float *dt, const float *R, const float *t)
{
  memcpy(dR, R, 9 * sizeof(float));
  if (dt && t)
    memcpy(dt, t, 3 * sizeof(float));
}

static void check_pair(const sage_ba_keyframe *a, const sage_ba_keyframe *b)
{
  SAGE_CHECK(a && b, "null keyframe");
  // the code size only matters for keyframes that carry a depth basis (frame 1 of the photometric / tracker operators has none)
  SAGE_CHECK(a->H == b->H && a->W == b->W && a->L == b->L && a->F == b->F && (a->C == b->C || !a->basis || !b->basis),
             "keyframes have different shapes");
}

// Tile-major sample order for the staged photometric kernels: samples sorted by (32 x 4 pixel tile, row in tile, column), so
// that the 4 consecutive batches of 32 samples a CTA works on cover one tile whose projection into the other frame is a small
// window.  Any sample set works (sub-sampled, masked); sparse ones just produce larger windows.  The order only changes the
// order of summation inside the kernels.  Thread-safe for concurrent problems sharing a keyframe.
void ensure_sorted_samples(sage_ba_context *ctx, sage_ba_keyframe *kf)
{
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (kf->sfeat_s || kf->N <= 0)
    return;
  SAGE_CHECK(kf->loc1d && kf->homo && kf->fg, "keyframe lacks sample / feature data");
  cudaStream_t s = ctx->stream;
  const int N = kf->N, W = kf->W;
  std::vector<int> loc(N), perm(N);
  SAGE_CUDA(cudaMemcpyAsync(loc.data(), kf->loc1d, sizeof(int) * N, cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  const int ntx = (W + 31) / 32;
  std::vector<uint64_t> key(N);
  for (int n = 0; n < N; ++n)
  {
    const int y = loc[n] / W, x = loc[n] - y * W;
    const uint64_t tile = (uint64_t)(y >> 2) * ntx + (x >> 5);
    key[n] = (((tile << 7) | (uint64_t)((y & 3) << 5) | (uint64_t)(x & 31)) << 32) | (uint32_t)n; // ties keep the caller's order
  }
  std::sort(key.begin(), key.end());
  for (int n = 0; n < N; ++n)
    perm[n] = (int)(key[n] & 0xffffffffu);
  int *dperm = nullptr;
  SAGE_CUDA(cudaMalloc(&dperm, sizeof(int) * N));
  int *loc_s = nullptr;
  float4 *homo_s = nullptr;
  float *sfeat_s = nullptr;
  try
  {
    SAGE_CUDA(cudaMemcpyAsync(dperm, perm.data(), sizeof(int) * N, cudaMemcpyHostToDevice, s));
    SAGE_CUDA(cudaMalloc(&loc_s, sizeof(int) * N));
    SAGE_CUDA(cudaMalloc(&homo_s, sizeof(float4) * N));
    SAGE_CUDA(cudaMalloc(&sfeat_s, sizeof(float) * (size_t)kf->L * N * kf->F));
    launch_permute_samples(dperm, kf->loc1d, kf->homo, loc_s, homo_s, N, s);
    launch_presample(kf->fg, nullptr, nullptr, loc_s, homo_s, nullptr, 1.f, kf->pyr, kf->F, kf->C, N, nullptr, nullptr, sfeat_s, s);
    ctx->launches += 2;
    SAGE_CUDA(cudaGetLastError());
    SAGE_CUDA(cudaStreamSynchronize(s));
  }
  catch (...)
  {
    cudaFree(dperm);
    cudaFree(loc_s);
    cudaFree(homo_s);
    cudaFree(sfeat_s);
    throw;
  }
  cudaFree(dperm);
  kf->loc1d_s = loc_s;
  kf->homo_s = homo_s;
  kf->sfeat_s = sfeat_s; // published last: the early-out above tests it
}

} // namespace sage

extern "C" {

const char *sage_ba_version(void) { return "sage-ba-b200 0.1 (sm_100a)"; }

int sage_ba_create(sage_ba_context **out, int device, void *stream)
{
  if (!out)
    return 1;
  *out = nullptr;
  sage_ba_context *ctx = new sage_ba_context();
  try
  {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw Error{std::string("no CUDA device available: ") + cudaGetErrorString(e)};
    SAGE_CHECK(device >= 0 && device < count, "device index out of range");
    ctx->device = device;
    SAGE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SAGE_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (stream)
      ctx->stream = (cudaStream_t)stream;
    else
    {
      SAGE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    SAGE_CHECK(cublasCreate(&ctx->cublas) == CUBLAS_STATUS_SUCCESS, "cublasCreate failed");
    SAGE_CHECK(cusolverDnCreate(&ctx->cusolver) == CUSOLVER_STATUS_SUCCESS, "cusolverDnCreate failed");
    cublasSetStream(ctx->cublas, ctx->stream);
    cusolverDnSetStream(ctx->cusolver, ctx->stream);
    *out = ctx;
    return 0;
  }
  catch (const Error &e)
  {
    fprintf(stderr, "sage_ba_create: %s\n", e.msg.c_str());
    delete ctx;
    return 1;
  }
}

void sage_ba_destroy(sage_ba_context *ctx)
{
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->cublas)
    cublasDestroy(ctx->cublas);
  if (ctx->cusolver)
    cusolverDnDestroy(ctx->cusolver);
  for (int q = 0; q < 2; ++q)
  {
    if (ctx->aux[q])
      cudaStreamDestroy(ctx->aux[q]);
    if (ctx->ev_join[q])
      cudaEventDestroy(ctx->ev_join[q]);
  }
  if (ctx->ev_fork)
    cudaEventDestroy(ctx->ev_fork);
  if (ctx->own_stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *sage_ba_last_error(const sage_ba_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
long sage_ba_launch_count(const sage_ba_context *ctx) { return ctx ? ctx->launches : 0; }
void *sage_ba_stream(const sage_ba_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int sage_ba_set_geometric_tcgen05(int on) { return sage::geo_set_tc(on); }

int sage_ba_synchronize(sage_ba_context *ctx)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaStreamSynchronize(ctx->stream));
  SAGE_CATCH
}

int sage_ba_keyframe_create(sage_ba_context *ctx, const sage_ba_keyframe_desc *d, sage_ba_keyframe **out)
{
  SAGE_TRY(ctx)
  SAGE_CHECK(d && out, "null argument");
  SAGE_CHECK(d->levels >= 1 && d->levels <= SAGE_BA_MAX_LEVELS, "levels out of range");
  SAGE_CHECK(d->feat_channels == 16 || d->feat_channels == 32, "feat_channels must be 16 or 32");
  SAGE_CHECK(d->code_size == 8 || d->code_size == 16 || d->code_size == 32, "code_size must be 8, 16 or 32");
  SAGE_CHECK((int)d->camera.width == d->width && (int)d->camera.height == d->height, "camera size does not match the maps");
  SAGE_CUDA(cudaSetDevice(ctx->device));
  sage_ba_keyframe *kf = new sage_ba_keyframe();
  kf->H = d->height;
  kf->W = d->width;
  kf->L = d->levels;
  kf->F = d->feat_channels;
  kf->C = d->code_size;
  kf->N = d->num_samples;
  kf->pyr = make_campyr(d->camera, d->levels, kf->cams);
  kf->SP = kf->pyr.off[kf->L - 1] + (long)kf->pyr.w[kf->L - 1] * kf->pyr.h[kf->L - 1];
  const long SP = kf->SP, HW = (long)kf->H * kf->W;
  const int F = kf->F, C = kf->C, N = kf->N;
  cudaStream_t s = ctx->stream;
  const bool host = d->memory == SAGE_BA_HOST;
  std::vector<void *> temps;
  int *bad_locs = nullptr;
  auto stage = [&](const void *src, size_t bytes) -> const void * {
    if (!host || !src)
      return src;
    void *t = nullptr;
    SAGE_CUDA(cudaMalloc(&t, bytes));
    temps.push_back(t);
    SAGE_CUDA(cudaMemcpyAsync(t, src, bytes, cudaMemcpyHostToDevice, s));
    return t;
  };
  try
  {
    // every group of tensors is optional (the reference's operators each see only part of a frame, and the df:: shim of
    // INTEGRATION.md builds partial keyframes from what one call is given); the entry points check for what they need
    SAGE_CHECK(d->feat_map_pyramid || d->feat_map || d->dpt_map_bias, "a keyframe needs feature maps or depth data");
    SAGE_CHECK(!d->feat_map || d->video_mask, "building the pyramid on the device needs video_mask");
    if (d->borrow_depth)
    {
      SAGE_CHECK(!host && d->dpt_map_bias && d->dpt_jac_code && d->video_mask && d->jac_stride_row == C && d->jac_stride_col == 1 &&
                     !d->feat_map && !d->feat_map_pyramid && N == 0,
                 "borrow_depth needs device bias / pixel-major basis / mask and nothing else");
      kf->bias = const_cast<float *>(d->dpt_map_bias);
      kf->basis = const_cast<float *>(d->dpt_jac_code);
      kf->mask = const_cast<float *>(d->video_mask);
      kf->borrowed_depth = true;
      *out = kf;
      return 0;
    }
    if (d->feat_map_pyramid || d->feat_map)
      SAGE_CUDA(cudaMalloc(&kf->fg, sizeof(float) * SP * 3 * F));
    if (d->video_mask)
      SAGE_CUDA(cudaMalloc(&kf->mask, sizeof(float) * HW));
    if (d->feat_map)
    {
      // device-side input builder: masked Gaussian pyramid + gradients straight into the channel-last layout
      const float *feat0 = (const float *)stage(d->feat_map, sizeof(float) * F * HW);
      const float *mask0 = (const float *)stage(d->video_mask, sizeof(float) * HW);
      void *mscr = nullptr;
      SAGE_CUDA(cudaMalloc(&mscr, sizeof(float) * (HW + 256)));
      temps.push_back(mscr);
      ctx->launches += launch_build_pyramid(feat0, mask0, kf->fg, (float *)mscr, kf->pyr, F, s);
    }
    else if (d->feat_map_pyramid)
    {
      const float *feat = (const float *)stage(d->feat_map_pyramid, sizeof(float) * F * SP);
      if (d->feat_map_grad_pyramid)
      {
        const float *grad = (const float *)stage(d->feat_map_grad_pyramid, sizeof(float) * 2 * F * SP);
        launch_relayout_fg(feat, grad, kf->fg, F, SP, s);
        ctx->launches += 3;
      }
      else
      {
        SAGE_CUDA(cudaMemsetAsync(kf->fg, 0, sizeof(float) * SP * 3 * F, s));
        launch_relayout_fg(feat, nullptr, kf->fg, F, SP, s);
        ctx->launches += 1;
      }
    }
    if (d->video_mask)
      SAGE_CUDA(cudaMemcpyAsync(kf->mask, d->video_mask, sizeof(float) * HW, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    if (d->dpt_map_bias)
    {
      SAGE_CHECK(d->dpt_jac_code, "dpt_jac_code is required with dpt_map_bias");
      SAGE_CUDA(cudaMalloc(&kf->bias, sizeof(float) * HW));
      SAGE_CUDA(cudaMalloc(&kf->basis, sizeof(float) * HW * C));
      SAGE_CUDA(cudaMalloc(&kf->dgm, sizeof(float4) * HW));
      SAGE_CUDA(cudaMalloc(&kf->dscr, sizeof(float) * HW));
      SAGE_CUDA(cudaMemcpyAsync(kf->bias, d->dpt_map_bias, sizeof(float) * HW, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
      // the strided view may address any element of the C*HW block
      const float *jac = (const float *)stage(d->dpt_jac_code, sizeof(float) * HW * C);
      launch_relayout_basis(jac, d->jac_stride_row, d->jac_stride_col, kf->basis, (int)HW, C, s);
      ctx->launches += 1;
    }
    if (N > 0)
    {
      SAGE_CHECK(d->sampled_locations_homo, "sampled_locations_homo is required when num_samples > 0");
      SAGE_CUDA(cudaMalloc(&kf->homo, sizeof(float4) * N));
      const float *h3 = (const float *)stage(d->sampled_locations_homo, sizeof(float) * 3 * N);
      launch_pack_homo(h3, kf->homo, N, s);
      ctx->launches += 1;
      if (d->sampled_locations_1d)
      {
        SAGE_CUDA(cudaMalloc(&kf->loc1d, sizeof(int) * N));
        const int64_t *l64 = (const int64_t *)stage(d->sampled_locations_1d, sizeof(int64_t) * N);
        SAGE_CUDA(cudaMalloc(&bad_locs, sizeof(int)));
        temps.push_back(bad_locs);
        SAGE_CUDA(cudaMemsetAsync(bad_locs, 0, sizeof(int), s));
        launch_convert_loc(l64, kf->loc1d, N, (int)HW, bad_locs, s);
        ctx->launches += 1;
        if (kf->fg)
        {
          // features of this keyframe at its own sample points, every level: what the mapping kernels read as "feat_0"
          SAGE_CUDA(cudaMalloc(&kf->sfeat, sizeof(float) * (size_t)kf->L * N * F));
          launch_presample(kf->fg, nullptr, nullptr, kf->loc1d, kf->homo, nullptr, 1.f, kf->pyr, F, C, N, nullptr, nullptr, kf->sfeat, s);
          ctx->launches += 1;
        }
      }
    }
    SAGE_CUDA(cudaGetLastError());
    SAGE_CUDA(cudaStreamSynchronize(s));
    if (bad_locs)
    {
      int nbad = 0;
      SAGE_CUDA(cudaMemcpy(&nbad, bad_locs, sizeof(int), cudaMemcpyDeviceToHost));
      SAGE_CHECK(nbad == 0, "sampled_locations_1d holds indices outside the image");
    }
  }
  catch (...)
  {
    for (void *t : temps)
      cudaFree(t);
    sage_ba_keyframe_destroy(ctx, kf);
    throw;
  }
  for (void *t : temps)
    cudaFree(t);
  *out = kf;
  SAGE_CATCH
}

int sage_ba_keyframe_set_bias(sage_ba_context *ctx, sage_ba_keyframe *kf, const float *dpt_map_bias, int memory)
{
  SAGE_TRY(ctx)
  SAGE_CHECK(kf && kf->bias && dpt_map_bias, "keyframe has no depth data");
  SAGE_CHECK(memory == SAGE_BA_HOST || memory == SAGE_BA_DEVICE, "memory must be SAGE_BA_HOST or SAGE_BA_DEVICE");
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CUDA(cudaMemcpyAsync(kf->bias, dpt_map_bias, sizeof(float) * kf->H * kf->W,
                            memory == SAGE_BA_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
  SAGE_CUDA(cudaStreamSynchronize(ctx->stream));
  SAGE_CATCH
}

void sage_ba_keyframe_destroy(sage_ba_context *ctx, sage_ba_keyframe *kf)
{
  if (!kf)
    return;
  if (ctx)
    cudaSetDevice(ctx->device);
  cudaFree(kf->fg);
  if (!kf->borrowed_depth)
  {
    cudaFree(kf->bias);
    cudaFree(kf->basis);
    cudaFree(kf->mask);
  }
  cudaFree(kf->loc1d);
  cudaFree(kf->homo);
  cudaFree(kf->sfeat);
  cudaFree(kf->dgm);
  cudaFree(kf->dscr);
  cudaFree(kf->loc1d_s);
  cudaFree(kf->homo_s);
  cudaFree(kf->sfeat_s);
  delete kf;
}

int sage_ba_keyframe_cameras(const sage_ba_keyframe *kf, sage_ba_camera *cams, int *level_offsets)
{
  if (!kf)
    return 1;
  for (int i = 0; i < kf->L; ++i)
  {
    if (cams)
      cams[i] = kf->cams[i];
    if (level_offsets)
      level_offsets[i] = kf->pyr.off[i];
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// single-factor operator API
// ------------------------------------------------------------------------------------------------
static PhotoFactor photo_map_factor(const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10, const float *t10,
                                    const float *R0, const float *t0, const float *R1, const float *code0, float scale0, float eps,
                                    const float *weights)
{
  check_pair(kf0, kf1);
  SAGE_CHECK(kf0->bias && kf0->basis && kf0->loc1d && kf0->homo && kf0->sfeat, "kf0 lacks depth / sample data");
  SAGE_CHECK(kf1->fg && kf1->mask, "kf1 lacks feature maps / mask");
  PhotoFactor f;
  memset(&f, 0, sizeof(f));
  f.fg0 = kf0->fg;
  f.sfeat0 = kf0->sfeat;
  f.fg1 = kf1->fg;
  f.mask1 = kf1->mask;
  f.bias0 = kf0->bias;
  f.basis0 = kf0->basis;
  f.loc1d = kf0->loc1d;
  f.homo = kf0->homo;
  f.N = kf0->N;
  fill_pose(f.R10, f.t10, R10, t10);
  if (R0)
    fill_pose(f.R0, f.t0, R0, t0);
  if (R1)
    fill_pose(f.R1, nullptr, R1, nullptr);
  memcpy(f.code0, code0, sizeof(float) * kf0->C);
  f.scale0 = scale0;
  f.eps = eps;
  memcpy(f.w, weights, sizeof(float) * kf0->L);
  return f;
}

int sage_ba_photometric_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10,
                                  const float *t10, const float *R0, const float *t0, const float *R1, const float *t1,
                                  const float *code0, float scale0, float eps, const float *weights, float *AtA, float *Atb,
                                  float *error, float *n_inliers)
{
  (void)t1;
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  PhotoFactor f = photo_map_factor(kf0, kf1, R10, t10, R0, t0, R1, code0, scale0, eps, weights);
  run_photo_single(ctx, PH_MAP_JAC, kf0->F, kf0->C, f, kf1->pyr, 13 + kf0->C, AtA, Atb, error, n_inliers);
  SAGE_CATCH
}

int sage_ba_photometric_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10,
                              const float *t10, const float *code0, float scale0, float eps, const float *weights, float *error,
                              float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  PhotoFactor f = photo_map_factor(kf0, kf1, R10, t10, nullptr, nullptr, nullptr, code0, scale0, eps, weights);
  run_photo_single(ctx, PH_MAP_ERR, kf0->F, kf0->C, f, kf1->pyr, 0, nullptr, nullptr, error, n_inliers);
  SAGE_CATCH
}

static PhotoFactor photo_trk_factor(const sage_ba_keyframe *fr1, const float *R, const float *t, const float *dpts, const float *homo4,
                                    const float *feats, int N, float scale0, float eps, const float *weights)
{
  SAGE_CHECK(fr1 && fr1->fg && fr1->mask, "frame lacks feature maps / mask");
  PhotoFactor f;
  memset(&f, 0, sizeof(f));
  f.fg1 = fr1->fg;
  f.mask1 = fr1->mask;
  f.homo = reinterpret_cast<const float4 *>(homo4);
  f.sfeat0 = feats;
  f.dpts0 = dpts;
  f.N = N;
  fill_pose(f.R10, f.t10, R, t);
  f.scale0 = scale0;
  f.eps = eps;
  f.dmul = 1.f;
  memcpy(f.w, weights, sizeof(float) * fr1->L);
  return f;
}

// homo [N,3] (device) -> packed float4 in the context scratch
static const float *pack_homo_scratch(sage_ba_context *ctx, const float *homo3, int N)
{
  float *h4 = ctx->scratch.ensure((size_t)N * 4);
  launch_pack_homo(homo3, reinterpret_cast<float4 *>(h4), N, ctx->stream);
  ctx->launches += 1;
  return h4;
}

} // extern "C"

namespace sage
{
// tracker photometric term with an explicit depth multiplier (TrackFrame optimises the scale of the sampled depths)
void run_tracker_photo(sage_ba_context *ctx, bool jac, const sage_ba_keyframe *frame1, const float *R, const float *t,
                       const float *dpts_dev, const float *homo3_dev, const float *feats_dev, int N, float dmul, float scale0, float eps,
                       const float *weights, float *AtA, float *Atb, float *error, float *n_inl)
{
  const float *h4 = pack_homo_scratch(ctx, homo3_dev, N);
  PhotoFactor f = photo_trk_factor(frame1, R, t, dpts_dev, h4, feats_dev, N, scale0, eps, weights);
  f.dmul = dmul;
  run_photo_single(ctx, jac ? PH_TRK_JAC : PH_TRK_ERR, frame1->F, frame1->C, f, frame1->pyr, scale0 != 0.f ? 7 : 6, AtA, Atb, error, n_inl);
}
} // namespace sage

extern "C" {

int sage_ba_tracker_photo_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *frame1, const float *R, const float *t,
                                    const float *sampled_dpts_0, const float *sampled_locations_homo_0, const float *sampled_features_0,
                                    int num_samples, int with_scale, float scale0, float eps, const float *weights, float *AtA,
                                    float *Atb, float *error, float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(!with_scale || scale0 != 0.f, "scale0 must be non-zero");
  const float *h4 = pack_homo_scratch(ctx, sampled_locations_homo_0, num_samples);
  PhotoFactor f = photo_trk_factor(frame1, R, t, sampled_dpts_0, h4, sampled_features_0, num_samples, with_scale ? scale0 : 0.f, eps, weights);
  run_photo_single(ctx, PH_TRK_JAC, frame1->F, frame1->C, f, frame1->pyr, with_scale ? 7 : 6, AtA, Atb, error, n_inliers);
  SAGE_CATCH
}

int sage_ba_tracker_photo_error(sage_ba_context *ctx, const sage_ba_keyframe *frame1, const float *R, const float *t,
                                const float *sampled_dpts_0, const float *sampled_locations_homo_0, const float *sampled_features_0,
                                int num_samples, float eps, const float *weights, float *error, float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  const float *h4 = pack_homo_scratch(ctx, sampled_locations_homo_0, num_samples);
  PhotoFactor f = photo_trk_factor(frame1, R, t, sampled_dpts_0, h4, sampled_features_0, num_samples, 0.f, eps, weights);
  run_photo_single(ctx, PH_TRK_ERR, frame1->F, frame1->C, f, frame1->pyr, 0, nullptr, nullptr, error, n_inliers);
  SAGE_CATCH
}

int sage_ba_tracker_presample(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const float *code0, float scale0, float *out_dpts,
                              float *out_homo, float *out_feats)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(kf0 && kf0->bias && kf0->loc1d && kf0->fg, "kf0 lacks depth / sample / feature data");
  float *dcode = ctx->tmp_code.ensure(SAGE_MAX_CODE);
  float *hc = reinterpret_cast<float *>(ctx->hfactor.ensure(1024));
  memcpy(hc, code0, sizeof(float) * kf0->C);
  SAGE_CUDA(cudaMemcpyAsync(dcode, hc, sizeof(float) * kf0->C, cudaMemcpyHostToDevice, ctx->stream));
  launch_presample(kf0->fg, kf0->bias, kf0->basis, kf0->loc1d, kf0->homo, dcode, scale0, kf0->pyr, kf0->F, kf0->C, kf0->N, out_dpts,
                   out_homo, out_feats, ctx->stream);
  ctx->launches += 1;
  SAGE_CUDA(cudaGetLastError());
  SAGE_CUDA(cudaStreamSynchronize(ctx->stream));
  SAGE_CATCH
}

// geometric ---------------------------------------------------------------------------------------
static void run_geo_single(sage_ba_context *ctx, bool jac, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10,
                           const float *t10, const float *R0, const float *t0, const float *R1, const float *code0, const float *code1,
                           float scale0, float scale1, float eps, float loss_param, float weight, float *AtA, float *Atb, float *error,
                           float *n_inl)
{
  check_pair(kf0, kf1);
  SAGE_CHECK(kf0->bias && kf0->loc1d && kf1->bias && kf1->mask, "keyframes lack depth / sample / mask data");
  const int C = kf0->C, D = 14 + 2 * C;
  cudaStream_t s = ctx->stream;
  // KF1's depth map + gradient from (code1) -- gtsam/geometric_factor.cpp:317-320 moved inside
  float *dcode = ctx->tmp_code.ensure(SAGE_MAX_CODE);
  float *hc = reinterpret_cast<float *>(ctx->hfactor.ensure(1024));
  memcpy(hc + 512 / 4, code1, sizeof(float) * C);
  SAGE_CUDA(cudaMemcpyAsync(dcode, hc + 512 / 4, sizeof(float) * C, cudaMemcpyHostToDevice, s));
  // KF1's state-dependent depth map lives in the CONTEXT, not in the keyframe: the reference calls this operator from several
  // threads that may share kf1 (TBB workers linearising new factors, NonlinearFactorGraph.cpp:329-335)
  const size_t HW1 = (size_t)kf1->H * kf1->W;
  float4 *dgm = reinterpret_cast<float4 *>(ctx->geo_dgm.ensure(4 * HW1));
  float *dscr = ctx->geo_dscr.ensure(HW1);
  launch_depth_maps(kf1->bias, kf1->basis, dcode, kf1->mask, dgm, dscr, kf1->H, kf1->W, C, s);
  ctx->launches += 2;

  GeoFactor f;
  memset(&f, 0, sizeof(f));
  f.bias0 = kf0->bias;
  f.basis0 = kf0->basis;
  f.loc1d = kf0->loc1d;
  f.homo = kf0->homo;
  f.dgm1 = dgm;
  f.basis1 = kf1->basis;
  f.N = kf0->N;
  fill_pose(f.R10, f.t10, R10, t10);
  if (R0)
    fill_pose(f.R0, f.t0, R0, t0);
  if (R1)
    fill_pose(f.R1, nullptr, R1, nullptr);
  memcpy(f.code0, code0, sizeof(float) * C);
  f.scale0 = scale0;
  f.scale1 = scale1;
  f.dscale = scale1;
  f.eps = eps;
  f.loss_param = loss_param;
  f.weight = weight;
  f.out = 0;

  const int WP = geo_row_width(C);
  const int sps = (32 / (C / 4)) * (SAGE_CTA / 32);
  const int slices = pick_slices(ctx, f.N, sps);
  const size_t nout = (jac ? (size_t)D * D + D : 0) + 2;
  const bool tc = geo_uses_tc(jac, C);
  float *partH = ctx->partH.ensure((size_t)slices * geo_partial_floats(jac, C, tc));
  float *partE = ctx->partE.ensure((size_t)slices * 2);
  float *out = ctx->out.ensure(nout);
  GeoFactor *df = reinterpret_cast<GeoFactor *>(ctx->factor.ensure(1024));
  GeoFactor *hf = reinterpret_cast<GeoFactor *>(ctx->hfactor.ensure(1024));
  static_assert(sizeof(GeoFactor) <= 512 && sizeof(PhotoFactor) <= 1024 && sizeof(ReprojFactor) <= 1024, "factor struct too large");
  *hf = f;
  SAGE_CUDA(cudaMemcpyAsync(df, hf, sizeof(GeoFactor), cudaMemcpyHostToDevice, s));
  const sage_ba_camera &cam = kf1->cams[0];
  SAGE_CHECK(launch_geo(jac, C, df, 1, kf1->W, kf1->H, cam.fx, cam.fy, cam.u0, cam.v0, slices, partH, partE, out, 1, s, tc) == 0,
             "unsupported code_size");
  ctx->launches += 2;
  SAGE_CUDA(cudaGetLastError());
  float *h = ctx->hout.ensure(nout);
  SAGE_CUDA(cudaMemcpyAsync(h, out, nout * sizeof(float), cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  if (jac)
  {
    memcpy(AtA, h, sizeof(float) * D * D);
    memcpy(Atb, h + D * D, sizeof(float) * D);
  }
  *error = h[nout - 2];
  if (n_inl)
    *n_inl = h[nout - 1];
}

int sage_ba_geometric_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10,
                                const float *t10, const float *R0, const float *t0, const float *R1, const float *t1,
                                const float *code0, const float *code1, float scale0, float scale1, float eps, float loss_param,
                                float weight, float *AtA, float *Atb, float *error, float *n_inliers)
{
  (void)t1;
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  run_geo_single(ctx, true, kf0, kf1, R10, t10, R0, t0, R1, code0, code1, scale0, scale1, eps, loss_param, weight, AtA, Atb, error,
                 n_inliers);
  SAGE_CATCH
}

int sage_ba_geometric_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10,
                            const float *t10, const float *code0, const float *code1, float scale0, float scale1, float eps,
                            float loss_param, float weight, float *error, float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  run_geo_single(ctx, false, kf0, kf1, R10, t10, nullptr, nullptr, nullptr, code0, code1, scale0, scale1, eps, loss_param, weight,
                 nullptr, nullptr, error, n_inliers);
  SAGE_CATCH
}

// reprojection ------------------------------------------------------------------------------------
static void run_reproj_single(sage_ba_context *ctx, bool jac, bool tracker, const sage_ba_keyframe *kf0, const sage_ba_camera &cam,
                              const float *R10, const float *t10, const float *R0, const float *t0, const float *R1, const float *code0,
                              float scale0, const int32_t *loc1d, const float *dpts, const float *homo, const float *match2d, int M,
                              float eps, float loss_param, float weight, float *AtA, float *Atb, float *error, float *n_inl)
{
  SAGE_CHECK(M >= 0 && M <= 4096, "num_matches out of range");
  if (!tracker)
    for (int m = 0; m < M; ++m) // bias / basis are gathered at these indices
      SAGE_CHECK(loc1d[m] >= 0 && loc1d[m] < kf0->H * kf0->W, "matched location outside the image");
  cudaStream_t s = ctx->stream;
  const int C = tracker ? 8 : kf0->C;
  const int D = tracker ? 6 : 13 + C;
  // upload the match arrays (they live on the host in this entry point)
  float *dm = ctx->trk_m_homo.ensure((size_t)std::max(M, 1) * 8);
  float *hm = ctx->hout.ensure(std::max<size_t>((size_t)M * 8, (size_t)D * D + D + 2));
  // layout: homo [M,3] | match2d [M,2] | dpts-or-loc [M] (as raw 32-bit)
  memcpy(hm, homo, sizeof(float) * 3 * M);
  memcpy(hm + 3 * M, match2d, sizeof(float) * 2 * M);
  if (tracker)
    memcpy(hm + 5 * M, dpts, sizeof(float) * M);
  else
    memcpy(hm + 5 * M, loc1d, sizeof(int32_t) * M);
  SAGE_CUDA(cudaMemcpyAsync(dm, hm, sizeof(float) * 6 * M, cudaMemcpyHostToDevice, s));
  SAGE_CUDA(cudaStreamSynchronize(s)); // hm is reused for the result below

  ReprojFactor f;
  memset(&f, 0, sizeof(f));
  f.homo = dm;
  f.match2d = dm + 3 * M;
  if (tracker)
    f.dpts0 = dm + 5 * M;
  else
  {
    SAGE_CHECK(kf0 && kf0->bias, "kf0 lacks depth data");
    f.loc1d = reinterpret_cast<const int *>(dm + 5 * M);
    f.bias0 = kf0->bias;
    f.basis0 = kf0->basis;
    memcpy(f.code0, code0, sizeof(float) * C);
  }
  f.M = M;
  fill_pose(f.R10, f.t10, R10, t10);
  if (R0)
    fill_pose(f.R0, f.t0, R0, t0);
  if (R1)
    fill_pose(f.R1, nullptr, R1, nullptr);
  f.scale0 = scale0;
  f.eps = eps;
  f.loss_param = loss_param;
  f.weight = weight;
  f.fx = cam.fx;
  f.fy = cam.fy;
  f.cx = cam.u0;
  f.cy = cam.v0;
  f.out = 0;
  const size_t nout = (jac ? (size_t)D * D + D : 0) + 2;
  float *out = ctx->out.ensure(nout);
  ReprojFactor *df = reinterpret_cast<ReprojFactor *>(ctx->factor.ensure(1024));
  ReprojFactor *hf = reinterpret_cast<ReprojFactor *>(ctx->hfactor.ensure(1024));
  *hf = f;
  SAGE_CUDA(cudaMemcpyAsync(df, hf, sizeof(ReprojFactor), cudaMemcpyHostToDevice, s));
  SAGE_CHECK(launch_reproj(jac, tracker, C, df, 1, out, 1, s) == 0, "unsupported code_size");
  ctx->launches += 1;
  SAGE_CUDA(cudaGetLastError());
  SAGE_CUDA(cudaMemcpyAsync(hm, out, nout * sizeof(float), cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  if (jac)
  {
    memcpy(AtA, hm, sizeof(float) * D * D);
    memcpy(Atb, hm + D * D, sizeof(float) * D);
  }
  *error = hm[nout - 2];
  if (n_inl)
    *n_inl = hm[nout - 1];
}

int sage_ba_reprojection_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const float *R10, const float *t10,
                                   const float *R0, const float *t0, const float *R1, const float *t1, const float *code0,
                                   float scale0, const int32_t *loc1d, const float *homo, const float *match2d, int num_matches,
                                   float eps, float loss_param, float weight, float *AtA, float *Atb, float *error, float *n_inliers)
{
  (void)t1;
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(kf0, "null keyframe");
  run_reproj_single(ctx, true, false, kf0, kf0->cams[0], R10, t10, R0, t0, R1, code0, scale0, loc1d, nullptr, homo, match2d, num_matches,
                    eps, loss_param, weight, AtA, Atb, error, n_inliers);
  SAGE_CATCH
}

int sage_ba_reprojection_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const float *R10, const float *t10, const float *code0,
                               float scale0, const int32_t *loc1d, const float *homo, const float *match2d, int num_matches, float eps,
                               float loss_param, float weight, float *error, float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(kf0, "null keyframe");
  run_reproj_single(ctx, false, false, kf0, kf0->cams[0], R10, t10, nullptr, nullptr, nullptr, code0, scale0, loc1d, nullptr, homo,
                    match2d, num_matches, eps, loss_param, weight, nullptr, nullptr, error, n_inliers);
  SAGE_CATCH
}

int sage_ba_tracker_reproj_jac_error(sage_ba_context *ctx, const sage_ba_camera *camera, const float *R, const float *t,
                                     const float *dpts, const float *homo, const float *match2d, int num_matches, float eps,
                                     float loss_param, float weight, float *AtA, float *Atb, float *error, float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(camera, "null camera");
  run_reproj_single(ctx, true, true, nullptr, *camera, R, t, nullptr, nullptr, nullptr, nullptr, 1.f, nullptr, dpts, homo, match2d,
                    num_matches, eps, loss_param, weight, AtA, Atb, error, n_inliers);
  SAGE_CATCH
}

int sage_ba_tracker_reproj_error(sage_ba_context *ctx, const sage_ba_camera *camera, const float *R, const float *t, const float *dpts,
                                 const float *homo, const float *match2d, int num_matches, float eps, float loss_param, float weight,
                                 float *error, float *n_inliers)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(camera, "null camera");
  run_reproj_single(ctx, false, true, nullptr, *camera, R, t, nullptr, nullptr, nullptr, nullptr, 1.f, nullptr, dpts, homo, match2d,
                    num_matches, eps, loss_param, weight, nullptr, nullptr, error, n_inliers);
  SAGE_CATCH
}

} // extern "C"

// tracker match-geometry ----------------------------------------------------------------------------
namespace sage
{
void run_match_geom_single(sage_ba_context *ctx, bool jac, const float *R, const float *t, const float *dpts0, const float *dpts1,
                           const float *homo0, const float *homo1, int M, float dmul, float scale0, float loss_param, float weight,
                           float *AtA, float *Atb, float *error)
{
  SAGE_CHECK(M >= 0 && M <= 4096, "num_matches out of range");
  cudaStream_t s = ctx->stream;
  const int D = scale0 != 0.f ? 7 : 6;
  float *dm = ctx->trk_m_homo.ensure((size_t)std::max(M, 1) * 8);
  float *hm = ctx->hout.ensure(std::max<size_t>((size_t)M * 8, (size_t)D * D + D + 2));
  memcpy(hm, homo0, sizeof(float) * 3 * M);
  memcpy(hm + 3 * M, homo1, sizeof(float) * 3 * M);
  memcpy(hm + 6 * M, dpts0, sizeof(float) * M);
  memcpy(hm + 7 * M, dpts1, sizeof(float) * M);
  SAGE_CUDA(cudaMemcpyAsync(dm, hm, sizeof(float) * 8 * M, cudaMemcpyHostToDevice, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  MatchGeomFactor f;
  memset(&f, 0, sizeof(f));
  f.homo0 = dm;
  f.homo1 = dm + 3 * M;
  f.dpts0 = dm + 6 * M;
  f.dpts1 = dm + 7 * M;
  f.M = M;
  memcpy(f.R, R, 9 * sizeof(float));
  memcpy(f.t, t, 3 * sizeof(float));
  f.dmul = dmul;
  f.scale0 = scale0;
  f.loss_param = loss_param;
  f.weight = weight;
  const size_t nout = (jac ? (size_t)D * D + D : 0) + 2;
  float *out = ctx->out.ensure(nout);
  MatchGeomFactor *df = reinterpret_cast<MatchGeomFactor *>(ctx->factor.ensure(1024));
  MatchGeomFactor *hf = reinterpret_cast<MatchGeomFactor *>(ctx->hfactor.ensure(1024));
  *hf = f;
  SAGE_CUDA(cudaMemcpyAsync(df, hf, sizeof(MatchGeomFactor), cudaMemcpyHostToDevice, s));
  launch_match_geom(jac, df, 1, out, 1, D, s);
  ctx->launches += 1;
  SAGE_CUDA(cudaGetLastError());
  SAGE_CUDA(cudaMemcpyAsync(hm, out, nout * sizeof(float), cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  if (jac)
  {
    memcpy(AtA, hm, sizeof(float) * D * D);
    memcpy(Atb, hm + D * D, sizeof(float) * D);
  }
  *error = hm[nout - 2];
}
} // namespace sage

extern "C" int sage_ba_tracker_match_geom_jac_error(sage_ba_context *ctx, const float *R, const float *t, const float *sampled_dpts_0,
                                                    const float *matched_dpts_1, const float *sampled_locations_homo_0,
                                                    const float *matched_locations_homo_1, int num_matches, int with_scale,
                                                    float scale0, float loss_param, float weight, float *AtA, float *Atb, float *error)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(!with_scale || scale0 != 0.f, "scale0 must be non-zero");
  run_match_geom_single(ctx, true, R, t, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0, matched_locations_homo_1, num_matches,
                        1.f, with_scale ? scale0 : 0.f, loss_param, weight, AtA, Atb, error);
  SAGE_CATCH
}

extern "C" int sage_ba_tracker_match_geom_error(sage_ba_context *ctx, const float *R, const float *t, const float *sampled_dpts_0,
                                                const float *matched_dpts_1, const float *sampled_locations_homo_0,
                                                const float *matched_locations_homo_1, int num_matches, float loss_param, float weight,
                                                float *error)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  run_match_geom_single(ctx, false, R, t, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0, matched_locations_homo_1, num_matches,
                        1.f, 0.f, loss_param, weight, nullptr, nullptr, error);
  SAGE_CATCH
}

// mapping-side match-geometry / loop-closure match-geometry ---------------------------------------------
namespace sage
{
static void run_map_match_geom_single(sage_ba_context *ctx, bool jac, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                                      const float *R10, const float *t10, const float *R0, const float *t0, const float *R1,
                                      const float *code0, const float *code1, float scale0, float scale1, const int32_t *loc0,
                                      const int32_t *loc1, const float *dpts0, const float *dpts1, const float *homo0, const float *homo1,
                                      int M, float loss_param, float weight, int loss_type, float *AtA, float *Atb, float *error)
{
  SAGE_CHECK(M >= 0 && M <= 4096, "num_matches out of range");
  SAGE_CHECK(loss_type >= 0 && loss_type <= 3, "unknown robust loss type");
  const bool loop = kf0 == nullptr;
  int C = 0;
  if (!loop)
  {
    SAGE_CHECK(kf1 && kf0->C == kf1->C, "keyframes disagree on code size");
    C = kf0->C;
    for (int m = 0; m < M; ++m)
      SAGE_CHECK(loc0[m] >= 0 && loc0[m] < kf0->H * kf0->W && loc1[m] >= 0 && loc1[m] < kf1->H * kf1->W, "match location out of range");
  }
  cudaStream_t s = ctx->stream;
  const int D = 14 + 2 * C;
  float *dm = ctx->trk_m_homo.ensure((size_t)std::max(M, 1) * 8);
  float *hm = ctx->hout.ensure(std::max<size_t>((size_t)M * 8, (size_t)D * D + D + 2));
  memcpy(hm, homo0, sizeof(float) * 3 * M);
  memcpy(hm + 3 * M, homo1, sizeof(float) * 3 * M);
  if (loop)
  {
    memcpy(hm + 6 * M, dpts0, sizeof(float) * M);
    memcpy(hm + 7 * M, dpts1, sizeof(float) * M);
  }
  else
  {
    memcpy(hm + 6 * M, loc0, sizeof(int32_t) * M);
    memcpy(hm + 7 * M, loc1, sizeof(int32_t) * M);
  }
  SAGE_CUDA(cudaMemcpyAsync(dm, hm, sizeof(float) * 8 * M, cudaMemcpyHostToDevice, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  MapMatchGeomFactor f;
  memset(&f, 0, sizeof(f));
  f.homo0 = dm;
  f.homo1 = dm + 3 * M;
  if (loop)
  {
    f.dpts0 = dm + 6 * M;
    f.dpts1 = dm + 7 * M;
  }
  else
  {
    f.loc0 = reinterpret_cast<const int *>(dm + 6 * M);
    f.loc1 = reinterpret_cast<const int *>(dm + 7 * M);
    f.bias0 = kf0->bias;
    f.basis0 = kf0->basis;
    f.bias1 = kf1->bias;
    f.basis1 = kf1->basis;
    memcpy(f.code0, code0, sizeof(float) * C);
    memcpy(f.code1, code1, sizeof(float) * C);
  }
  f.M = M;
  memcpy(f.R10, R10, 9 * sizeof(float));
  memcpy(f.t10, t10, 3 * sizeof(float));
  if (jac)
  {
    memcpy(f.R0, R0, 9 * sizeof(float));
    memcpy(f.t0, t0, 3 * sizeof(float));
    memcpy(f.R1, R1, 9 * sizeof(float));
  }
  f.scale0 = scale0;
  f.scale1 = scale1;
  f.loss_param = loss_param;
  f.weight = weight;
  f.loss_type = loss_type;
  const size_t nout = (jac ? (size_t)D * D + D : 0) + 2;
  float *out = ctx->out.ensure(nout);
  MapMatchGeomFactor *df = reinterpret_cast<MapMatchGeomFactor *>(ctx->factor.ensure(1024));
  MapMatchGeomFactor *hf = reinterpret_cast<MapMatchGeomFactor *>(ctx->hfactor.ensure(1024));
  static_assert(sizeof(MapMatchGeomFactor) <= 1024, "factor staging buffer too small");
  *hf = f;
  SAGE_CUDA(cudaMemcpyAsync(df, hf, sizeof(MapMatchGeomFactor), cudaMemcpyHostToDevice, s));
  SAGE_CHECK(launch_map_match_geom(jac, C, df, 1, out, 1, s) == 0, "unsupported code_size");
  ctx->launches += 1;
  SAGE_CUDA(cudaGetLastError());
  SAGE_CUDA(cudaMemcpyAsync(hm, out, nout * sizeof(float), cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  if (jac)
  {
    memcpy(AtA, hm, sizeof(float) * D * D);
    memcpy(Atb, hm + D * D, sizeof(float) * D);
  }
  *error = hm[nout - 2];
}
} // namespace sage

extern "C" int sage_ba_match_geometry_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                                                const float *R10, const float *t10, const float *R0, const float *t0, const float *R1,
                                                const float *t1, const float *code0, const float *code1, float scale0, float scale1,
                                                const int32_t *sampled_locations_1d_0, const int32_t *matched_locations_1d_1,
                                                const float *sampled_locations_homo_0, const float *matched_locations_homo_1,
                                                int num_matches, float loss_param, float weight, int loss_type, float *AtA, float *Atb,
                                                float *error)
{
  (void)t1;
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(kf0 && kf1, "null keyframe");
  SAGE_CHECK(scale0 != 0.f && scale1 != 0.f, "scales must be non-zero");
  run_map_match_geom_single(ctx, true, kf0, kf1, R10, t10, R0, t0, R1, code0, code1, scale0, scale1, sampled_locations_1d_0,
                            matched_locations_1d_1, nullptr, nullptr, sampled_locations_homo_0, matched_locations_homo_1, num_matches,
                            loss_param, weight, loss_type, AtA, Atb, error);
  SAGE_CATCH
}

extern "C" int sage_ba_match_geometry_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1, const float *R10,
                                            const float *t10, const float *code0, const float *code1, float scale0, float scale1,
                                            const int32_t *sampled_locations_1d_0, const int32_t *matched_locations_1d_1,
                                            const float *sampled_locations_homo_0, const float *matched_locations_homo_1, int num_matches,
                                            float loss_param, float weight, int loss_type, float *error)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  SAGE_CHECK(kf0 && kf1, "null keyframe");
  run_map_match_geom_single(ctx, false, kf0, kf1, R10, t10, nullptr, nullptr, nullptr, code0, code1, scale0, scale1,
                            sampled_locations_1d_0, matched_locations_1d_1, nullptr, nullptr, sampled_locations_homo_0,
                            matched_locations_homo_1, num_matches, loss_param, weight, loss_type, nullptr, nullptr, error);
  SAGE_CATCH
}

extern "C" int sage_ba_loop_mg_jac_error(sage_ba_context *ctx, const float *R10, const float *t10, const float *R0, const float *t0,
                                         const float *R1, const float *t1, const float *sampled_unscaled_dpts_0,
                                         const float *matched_unscaled_dpts_1, const float *sampled_locations_homo_0,
                                         const float *matched_locations_homo_1, int num_matches, float scale0, float scale1,
                                         float loss_param, float weight, float *AtA, float *Atb, float *error)
{
  (void)t1;
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  run_map_match_geom_single(ctx, true, nullptr, nullptr, R10, t10, R0, t0, R1, nullptr, nullptr, scale0, scale1, nullptr, nullptr,
                            sampled_unscaled_dpts_0, matched_unscaled_dpts_1, sampled_locations_homo_0, matched_locations_homo_1,
                            num_matches, loss_param, weight, 0, AtA, Atb, error);
  SAGE_CATCH
}

extern "C" int sage_ba_loop_mg_error(sage_ba_context *ctx, const float *R10, const float *t10, const float *sampled_unscaled_dpts_0,
                                     const float *matched_unscaled_dpts_1, const float *sampled_locations_homo_0,
                                     const float *matched_locations_homo_1, int num_matches, float scale0, float scale1, float loss_param,
                                     float weight, float *error)
{
  SAGE_TRY(ctx)
  SAGE_CUDA(cudaSetDevice(ctx->device));
  run_map_match_geom_single(ctx, false, nullptr, nullptr, R10, t10, nullptr, nullptr, nullptr, nullptr, nullptr, scale0, scale1, nullptr,
                            nullptr, sampled_unscaled_dpts_0, matched_unscaled_dpts_1, sampled_locations_homo_0,
                            matched_locations_homo_1, num_matches, loss_param, weight, 0, nullptr, nullptr, error);
  SAGE_CATCH
}

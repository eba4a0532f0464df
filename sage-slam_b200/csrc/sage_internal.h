// sage_internal.h -- host-side objects behind the opaque handles of include/sage_ba.h.
#pragma once
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sage_ba.h"
#include "sage_kernels.h"

namespace sage
{

struct Error
{
  std::string msg;
};

#define SAGE_CUDA(call)                                                                                   \
  do                                                                                                      \
  {                                                                                                       \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      throw sage::Error{std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +   \
                        std::to_string(__LINE__) + ")"};                                                  \
  } while (0)

#define SAGE_CHECK(cond, text)                \
  do                                          \
  {                                           \
    if (!(cond))                              \
      throw sage::Error{std::string(text)};   \
  } while (0)

template <typename T>
struct DevBuf
{
  T *p = nullptr;
  size_t cap = 0;
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  T *ensure(size_t n)
  {
    if (n > cap)
    {
      release();
      SAGE_CUDA(cudaMalloc(&p, n * sizeof(T)));
      cap = n;
    }
    return p;
  }
};

template <typename T>
struct PinBuf
{
  T *p = nullptr;
  size_t cap = 0;
  ~PinBuf()
  {
    if (p)
      cudaFreeHost(p);
  }
  T *ensure(size_t n)
  {
    if (n > cap)
    {
      if (p)
        cudaFreeHost(p);
      p = nullptr;
      SAGE_CUDA(cudaMallocHost(&p, n * sizeof(T)));
      cap = n;
    }
    return p;
  }
};

} // namespace sage

struct sage_ba_context;
struct sage_ba_keyframe;
namespace sage
{
void run_tracker_photo(sage_ba_context *ctx, bool jac, const sage_ba_keyframe *frame1, const float *R, const float *t,
                       const float *dpts_dev, const float *homo3_dev, const float *feats_dev, int N, float dmul, float scale0, float eps,
                       const float *weights, float *AtA, float *Atb, float *error, float *n_inl);
void run_match_geom_single(sage_ba_context *ctx, bool jac, const float *R, const float *t, const float *dpts0, const float *dpts1,
                           const float *homo0, const float *homo1, int M, float dmul, float scale0, float loss_param, float weight,
                           float *AtA, float *Atb, float *error);
} // namespace sage

struct sage_ba_context
{
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // side streams of the batched problem: the three factor kinds of an iteration are independent launches, forked off the main
  // stream and joined before the exchange / assembly (events without timing)
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  std::string err;
  long launches = 0;
  int num_sms = 148;
  // workspaces of the single-factor entry points
  sage::DevBuf<float> partH, partE, out, scratch;
  sage::DevBuf<unsigned char> factor; // one factor struct
  sage::PinBuf<float> hout;
  sage::PinBuf<unsigned char> hfactor;
  sage::DevBuf<float> tmp_code;
  sage::DevBuf<float> geo_dgm, geo_dscr; // KF1 depth map (D, dx, dy, mask) of the single-factor geometric operator
  // tracker scratch
  sage::DevBuf<float> trk_dpts, trk_homo, trk_feats, trk_m_dpts, trk_m_homo, trk_m_2d;
  // descriptor cycle-matching scratch (descriptor.cu)
  sage::DevBuf<float> dm_maps, dm_float;
  sage::DevBuf<int> dm_int;
  sage::PinBuf<int> dm_hint;
};

struct sage_ba_keyframe
{
  int H = 0, W = 0, L = 0, F = 0, C = 0, N = 0;
  long SP = 0;
  sage_ba_camera cams[SAGE_BA_MAX_LEVELS];
  CamPyr pyr;
  float *fg = nullptr;    // [SP][3][F]
  float *bias = nullptr;  // [HW]
  float *basis = nullptr; // [HW][C]
  float *mask = nullptr;  // [H][W]
  int *loc1d = nullptr;   // [N]
  float4 *homo = nullptr; // [N]
  float *sfeat = nullptr; // [L][N][F] features pre-sampled at the keyframe's own sample points (camera_tracker.cpp:1104-1123)
  float4 *dgm = nullptr;  // [HW] (D, dx, dy, mask) for the geometric factor, state dependent
  float *dscr = nullptr;  // [HW] scratch
  // the same samples in TILE-MAJOR order (32 x 4 pixel tiles, raster inside a tile) for the staged photometric kernels of the
  // batched problem; built on first use by sage::ensure_sorted_samples
  int *loc1d_s = nullptr;
  float4 *homo_s = nullptr;
  float *sfeat_s = nullptr;
  bool borrowed_depth = false; // bias / basis / mask alias the caller's device memory (sage_ba_keyframe_desc.borrow_depth)
};
namespace sage
{
void ensure_sorted_samples(sage_ba_context *ctx, sage_ba_keyframe *kf);
}

// Every extern "C" entry point wraps its body in SAGE_TRY(ctx) ... SAGE_CATCH: exceptions become a non-zero return code
// and the message is kept in the context for sage_ba_last_error().
#define SAGE_TRY(ctx_) \
  sage_ba_context *ctx__ = (ctx_); \
  try                  \
  {
#define SAGE_CATCH                                   \
  }                                                  \
  catch (const sage::Error &e)                       \
  {                                                  \
    if (ctx__)                                       \
      ctx__->err = e.msg;                            \
    return 1;                                        \
  }                                                  \
  catch (const std::exception &e)                    \
  {                                                  \
    if (ctx__)                                       \
      ctx__->err = e.what();                         \
    return 1;                                        \
  }                                                  \
  return 0;

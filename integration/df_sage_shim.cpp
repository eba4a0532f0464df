// df_sage_shim.cpp -- the reference's `df::*_calculate` symbols implemented on top of the C ABI of include/sage_ba.h.
//
// This is the binding INTEGRATION.md describes: a maintainer of the reference replaces the four translation units of the
// static library df_cuda (system/sources/cuda/CMakeLists.txt:37-46) by this file and links libsage_ba.so; the reference's
// headers (cuda/*_factor_kernels.h) stay untouched, so core/gtsam/*_factor.cpp and core/system/camera_tracker.cpp compile
// and link unchanged.  It is compiled here against those headers (oracle/build_ref.py, only where /root/reference exists)
// together with the same pybind front that drives the reference's own kernels, and tests/test_shim_dropin.py holds its
// outputs to the reference goldens -- the df:: boundary itself is under test, not only the C ABI below it.
//
// Every reference operator receives only a PART of each frame (the photometric one sees KF0's features, depth and samples
// and KF1's features, gradients and mask), so the shim builds partial sage_ba keyframes from the tensors of the call and
// caches them by the identity (data pointers) of those tensors; in the live system the cache is filled once per frame
// (INTEGRATION.md section 2).  CUDA errors keep the reference's behaviour: message on stderr and exit()
// (cuda/photometric_factor_kernels.cpp:23-31).
#include <c10/cuda/CUDAStream.h>
#include <torch/torch.h>

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "geometric_factor_kernels.h"
#include "match_geometry_factor_kernels.h"
#include "photometric_factor_kernels.h"
#include "reprojection_factor_kernels.h"
#include "sage_ba.h"

namespace
{

sage_ba_context *ctx()
{
  // one context per calling thread (the reference is entered from up to four, core/deepfactors.cpp:1497-1505),
  // running on torch's current stream so that the call is ordered after the producers of its input tensors
  thread_local sage_ba_context *c = nullptr;
  if (!c)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (sage_ba_create(&c, dev, c10::cuda::getCurrentCUDAStream(dev).stream()) != 0)
    {
      fprintf(stderr, "sage_ba_create failed\n");
      exit(1);
    }
  }
  // the reference's operators are synchronous; when torch's current stream is the legacy default stream the context runs
  // on a private stream, so wait here for whatever produced the input tensors
  cudaStreamSynchronize(c10::cuda::getCurrentCUDAStream().stream());
  return c;
}

#define SAGE_OK(call)                                                         \
  do                                                                          \
  {                                                                           \
    if ((call) != 0)                                                          \
    {                                                                         \
      fprintf(stderr, "sage_ba: %s (%s)\n", sage_ba_last_error(ctx()), #call); \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

// small state tensors (poses, codes, weights, match arrays) -> contiguous host float / int32
std::vector<float> hostf(const at::Tensor &t)
{
  const at::Tensor c = t.to(at::kCPU, at::kFloat).contiguous();
  return std::vector<float>(c.data_ptr<float>(), c.data_ptr<float>() + c.numel());
}
std::vector<int32_t> hosti(const at::Tensor &t)
{
  const at::Tensor c = t.to(at::kCPU, at::kInt).contiguous();
  return std::vector<int32_t>(c.data_ptr<int32_t>(), c.data_ptr<int32_t>() + c.numel());
}
// Several small DEVICE tensors in ONE device-to-host transfer (one concatenation kernel + one blocking copy instead of one of
// each per tensor): the reference's callers hand over 6-8 tensors of 3-32 floats per call (gtsam/photometric_factor.cpp:264-311).
struct HostPack
{
  std::vector<float> buf;
  std::vector<size_t> off;
  const float *operator[](size_t i) const { return buf.data() + off[i]; }
  size_t count(size_t i) const { return off[i + 1] - off[i]; }
};
HostPack hostpack(std::initializer_list<at::Tensor> ts)
{
  HostPack h;
  std::vector<at::Tensor> flat;
  size_t total = 0;
  h.off.push_back(0);
  bool all_cuda = true;
  for (const at::Tensor &t : ts)
  {
    all_cuda = all_cuda && t.is_cuda();
    total += (size_t)t.numel();
    h.off.push_back(total);
  }
  h.buf.resize(total);
  if (all_cuda && ts.size() > 1)
  {
    for (const at::Tensor &t : ts)
      flat.push_back(t.to(at::kFloat).reshape({-1}));
    const at::Tensor c = torch::cat(flat).to(at::kCPU);
    std::copy(c.data_ptr<float>(), c.data_ptr<float>() + total, h.buf.begin());
  }
  else
  {
    size_t i = 0;
    for (const at::Tensor &t : ts)
    {
      const at::Tensor c = t.to(at::kCPU, at::kFloat).contiguous();
      std::copy(c.data_ptr<float>(), c.data_ptr<float>() + c.numel(), h.buf.begin() + h.off[i++]);
    }
  }
  return h;
}

struct KfDeleter
{
  sage_ba_context *c;
  void operator()(sage_ba_keyframe *k) const { sage_ba_keyframe_destroy(c, k); }
};
using KfPtr = std::shared_ptr<sage_ba_keyframe>;
// identity of a cached keyframe: the frame's persistent tensors (data pointers) + shape.  The gradient pyramid is NOT part of
// it (an entry built without gradients is upgraded in place when a caller needs them: one entry per frame and role, whichever
// operator sees the frame first), and the code size only counts when the view carries a depth basis (the tracker and the
// photometric operators see frame 1 without one).
using KfKey = std::tuple<const void *, const void *, const void *, const void *, long, long, long, long, long, long>;
using WeakStorage = c10::weak_intrusive_ptr<c10::StorageImpl>;
struct KfEntry
{
  KfPtr kf;
  bool has_grad = false;
  // the caller's tensors by WEAK reference to their storage: the cache pins no memory, an entry is valid exactly as long as the
  // storage it was built from is alive, and an address the allocator hands out again can never alias a stale entry (the storage
  // object differs).  (Strong references would also defeat the liveness test under a Python host, where a tensor once seen by
  // Python keeps use_count() == 2 for as long as any C++ handle exists.)
  std::vector<WeakStorage> src;
  std::vector<const void *> src_id;
  uint64_t last_use = 0;
};
// Heap objects that are never destroyed: at process exit torch and the CUDA context may already be gone.
std::map<KfKey, KfEntry> &g_cache = *new std::map<KfKey, KfEntry>();
uint64_t g_tick = 0;
std::mutex g_mutex;

size_t cache_cap()
{
  static const size_t cap = [] {
    const char *e = getenv("SAGE_SHIM_CACHE_MAX");
    return e ? (size_t)std::max(4, atoi(e)) : (size_t)128;
  }();
  return cap;
}

// Drop the entries whose source tensors nobody but the cache references any more (the frame was released, or the tensor was a
// per-call temporary), then the least recently used ones beyond the cap.  Called with g_mutex held, on every cache miss.
void evict_dead_entries()
{
  for (auto it = g_cache.begin(); it != g_cache.end();)
  {
    bool dead = false;
    for (const WeakStorage &w : it->second.src)
      dead = dead || w.expired();
    if (dead)
      it = g_cache.erase(it); // a keyframe still in use by a running call lives on through its shared_ptr
    else
      ++it;
  }
  while (g_cache.size() >= cache_cap())
  {
    auto lru = g_cache.begin();
    for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
      if (it->second.last_use < lru->second.last_use)
        lru = it;
    g_cache.erase(lru);
  }
}

struct FrameView
{
  // any subset of a frame's tensors, as the reference operators receive them
  at::Tensor feat_pyramid, grad_pyramid, bias, jac, mask, loc1d, homo;
};

KfPtr keyframe(const FrameView &v, const df::PinholeCamera<float> &cam0, int levels, int F, int C, bool key_bias = true)
{
  auto ptr = [](const at::Tensor &t) -> const void * { return t.defined() ? t.data_ptr() : nullptr; };
  // the sample locations are a frame member too, but callers hand over per-call casts of them (geometric_factor.cpp:344), so
  // only their count takes part.  key_bias lets a caller key a keyframe whose bias is state dependent (KF1 of the geometric
  // operators) on its other tensors only.
  const KfKey key{ptr(v.feat_pyramid), key_bias ? ptr(v.bias) : nullptr, ptr(v.jac), ptr(v.mask),
                  v.homo.defined() ? (long)v.homo.size(0) : 0, (long)cam0.width(), (long)cam0.height(), (long)levels,
                  v.feat_pyramid.defined() ? (long)F : 0, v.jac.defined() ? (long)C : 0};
  const bool want_grad = v.grad_pyramid.defined();
  // storages the entry depends on (identity + liveness)
  std::vector<const void *> ids;
  std::vector<WeakStorage> weak;
  for (const at::Tensor *t : {&v.feat_pyramid, &v.jac, &v.mask, key_bias ? &v.bias : nullptr})
    if (t && t->defined())
    {
      ids.push_back(t->storage().unsafeGetStorageImpl());
      weak.push_back(t->storage().getWeakStorageImpl());
    }
  std::lock_guard<std::mutex> lock(g_mutex);
  {
    auto it = g_cache.find(key);
    if (it != g_cache.end())
    {
      bool same = it->second.src_id == ids;
      for (const WeakStorage &w : it->second.src)
        same = same && !w.expired();
      if (!same)
      {
        g_cache.erase(it); // the address was handed out again for other data: not the same frame
        it = g_cache.end();
      }
    }
    if (it != g_cache.end() && (it->second.has_grad || !want_grad))
    {
      it->second.last_use = ++g_tick;
      if (!key_bias) // same frame, new state: refresh the depth map in place
        SAGE_OK(sage_ba_keyframe_set_bias(ctx(), it->second.kf.get(), v.bias.to(at::kFloat).contiguous().data_ptr<float>(), SAGE_BA_DEVICE));
      return it->second.kf;
    }
    if (it != g_cache.end())
      g_cache.erase(it); // cached without gradients, needed with: rebuild below (the old copy dies with its last user)
  }
  evict_dead_entries();
  sage_ba_keyframe_desc d{};
  d.memory = SAGE_BA_DEVICE;
  d.height = (int)cam0.height();
  d.width = (int)cam0.width();
  d.levels = levels;
  d.feat_channels = F;
  d.code_size = C;
  d.camera = {cam0.fx(), cam0.fy(), cam0.u0(), cam0.v0(), cam0.width(), cam0.height()};
  std::vector<at::Tensor> keep; // contiguous float copies, only needed until sage_ba_keyframe_create returns (it synchronises)
  auto cf = [&](const at::Tensor &t) -> const float * {
    if (!t.defined())
      return nullptr;
    keep.push_back(t.to(at::kFloat).contiguous());
    return keep.back().data_ptr<float>();
  };
  d.feat_map_pyramid = cf(v.feat_pyramid);
  d.feat_map_grad_pyramid = cf(v.grad_pyramid);
  d.dpt_map_bias = cf(v.bias);
  d.video_mask = cf(v.mask);
  if (v.jac.defined())
  {
    // the depth basis arrives as a strided [HW, C] view of the network's [C, H, W] output (code_depth_network.cpp:38-39)
    d.dpt_jac_code = v.jac.data_ptr<float>();
    d.jac_stride_row = v.jac.stride(0);
    d.jac_stride_col = v.jac.stride(1);
  }
  if (v.homo.defined())
  {
    d.sampled_locations_homo = cf(v.homo);
    d.num_samples = (int)v.homo.size(0);
    if (v.loc1d.defined())
    {
      // int64 in the photometric path, already cast to int32 by the geometric / reprojection callers (geometric_factor.cpp:344)
      keep.push_back(v.loc1d.to(at::kLong).contiguous());
      d.sampled_locations_1d = keep.back().data_ptr<int64_t>();
    }
  }
  sage_ba_keyframe *kf = nullptr;
  sage_ba_context *c = ctx();
  SAGE_OK(sage_ba_keyframe_create(c, &d, &kf));
  KfEntry e;
  e.kf = KfPtr(kf, KfDeleter{c});
  e.has_grad = want_grad;
  e.last_use = ++g_tick;
  e.src = std::move(weak);
  e.src_id = std::move(ids);
  KfPtr out = e.kf;
  g_cache[key] = std::move(e);
  return out;
}

long cache_size_locked()
{
  std::lock_guard<std::mutex> lock(g_mutex);
  return (long)g_cache.size();
}

// KF1 of the geometric operators arrives as per-call tensors (a state-dependent depth map and a fresh pixel-major copy of the
// basis, gtsam/geometric_factor.cpp:340-342): a handle that BORROWS them (no copy, no allocation, not cached) lives for the call
KfPtr borrowed_depth_keyframe(const at::Tensor &bias, const at::Tensor &basis_hwc, const at::Tensor &mask, const df::PinholeCamera<float> &cam0,
                              int C)
{
  sage_ba_keyframe_desc d{};
  d.memory = SAGE_BA_DEVICE;
  d.height = (int)cam0.height();
  d.width = (int)cam0.width();
  d.levels = 1;
  d.feat_channels = 16;
  d.code_size = C;
  d.camera = {cam0.fx(), cam0.fy(), cam0.u0(), cam0.v0(), cam0.width(), cam0.height()};
  d.dpt_map_bias = bias.data_ptr<float>();
  d.dpt_jac_code = basis_hwc.data_ptr<float>();
  d.jac_stride_row = C;
  d.jac_stride_col = 1;
  d.video_mask = mask.data_ptr<float>();
  d.borrow_depth = 1;
  sage_ba_keyframe *kf = nullptr;
  sage_ba_context *c = ctx();
  SAGE_OK(sage_ba_keyframe_create(c, &d, &kf));
  return KfPtr(kf, KfDeleter{c});
}

at::Tensor to_dev(const float *p, std::vector<int64_t> shape, const at::Tensor &like)
{
  return torch::from_blob(const_cast<float *>(p), shape, at::kFloat).clone().to(like.device());
}

int loss_type(const std::string &s)
{
  if (s == "fair")
    return SAGE_BA_LOSS_FAIR;
  if (s == "L2")
    return SAGE_BA_LOSS_L2;
  if (s == "huber")
    return SAGE_BA_LOSS_HUBER;
  if (s == "unbiased")
    return SAGE_BA_LOSS_UNBIASED;
  fprintf(stderr, "sage_ba: unknown robust loss type %s\n", s.c_str());
  exit(1);
}

sage_ba_camera cam_of(const df::PinholeCamera<float> &c) { return {c.fx(), c.fy(), c.u0(), c.v0(), c.width(), c.height()}; }

} // namespace

// test hook: number of cached keyframes (tests/test_shim_dropin.py checks that the cache does not grow with released frames)
extern "C" __attribute__((visibility("default"))) long sage_shim_cache_size() { return cache_size_locked(); }

namespace df
{

// ------------------------------------------------------------------------------------------------ photometric
template <int FS>
float photometric_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor flatten_dpt_map_bias_0,
                                  const at::Tensor flatten_dpt_jac_code_0, const at::Tensor code_0, const at::Tensor valid_mask_1,
                                  const at::Tensor sampled_locations_1d_0, const at::Tensor sampled_locations_homo_0,
                                  const at::Tensor feat_map_pyramid_0, const at::Tensor feat_map_pyramid_1, const at::Tensor level_offsets,
                                  const float scale_0, const CameraPyramid<float> &camera_pyramid, const float eps,
                                  const at::Tensor weights_tensor)
{
  const int L = (int)level_offsets.size(0), C = (int)code_0.numel();
  KfPtr kf0 = keyframe({feat_map_pyramid_0, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, sampled_locations_1d_0,
                        sampled_locations_homo_0},
                       camera_pyramid[0], L, FS, C);
  KfPtr kf1 = keyframe({feat_map_pyramid_1, {}, {}, {}, valid_mask_1, {}, {}}, camera_pyramid[0], L, FS, C);
  const HostPack h = hostpack({rotation, translation, code_0});
  const auto w = hostf(weights_tensor); // a CPU tensor in the mapping path (photometric_factor.cpp:31-32)
  float err = 0.f;
  SAGE_OK(sage_ba_photometric_error(ctx(), kf0.get(), kf1.get(), h[0], h[1], h[2], scale_0, eps, w.data(), &err, nullptr));
  return err;
}

template <int CS, int FS>
void photometric_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation10,
                                     const at::Tensor translation10, const at::Tensor rotation0, const at::Tensor translation0,
                                     const at::Tensor rotation1, const at::Tensor translation1, const at::Tensor flatten_dpt_map_bias_0,
                                     const at::Tensor flatten_dpt_jac_code_0, const at::Tensor code_0, const at::Tensor valid_mask_1,
                                     const at::Tensor sampled_locations_1d_0, const at::Tensor sampled_locations_homo_0,
                                     const at::Tensor feat_map_pyramid_0, const at::Tensor feat_map_pyramid_1,
                                     const at::Tensor feat_map_grad_pyramid_1, const at::Tensor level_offsets, const float scale_0,
                                     const CameraPyramid<float> &camera_pyramid, const float eps, const at::Tensor weights_tensor)
{
  const int L = (int)level_offsets.size(0);
  KfPtr kf0 = keyframe({feat_map_pyramid_0, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, sampled_locations_1d_0,
                        sampled_locations_homo_0},
                       camera_pyramid[0], L, FS, CS);
  KfPtr kf1 = keyframe({feat_map_pyramid_1, feat_map_grad_pyramid_1, {}, {}, valid_mask_1, {}, {}}, camera_pyramid[0], L, FS, CS);
  constexpr int D = 13 + CS;
  std::vector<float> A(D * D), b(D);
  const HostPack h = hostpack({rotation10, translation10, rotation0, translation0, rotation1, translation1, code_0});
  const auto w = hostf(weights_tensor);
  SAGE_OK(sage_ba_photometric_jac_error(ctx(), kf0.get(), kf1.get(), h[0], h[1], h[2], h[3], h[4], h[5], h[6], scale_0, eps, w.data(),
                                        A.data(), b.data(), &error, nullptr));
  AtA = to_dev(A.data(), {D, D}, rotation10);
  Atb = to_dev(b.data(), {D, 1}, rotation10);
}

template <int FS>
static void tracker_photo_jac(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor &rotation, const at::Tensor &translation,
                              const at::Tensor &valid_mask_1, const at::Tensor &sampled_dpts_0, const at::Tensor &sampled_locations_homo_0,
                              const at::Tensor &sampled_features_0, const at::Tensor &feat_map_pyramid_1,
                              const at::Tensor &feat_map_grad_pyramid_1, const at::Tensor &level_offsets,
                              const CameraPyramid<float> &camera_pyramid, int with_scale, float scale_0, float eps,
                              const at::Tensor &weights_tensor)
{
  const int L = (int)level_offsets.size(0), D = with_scale ? 7 : 6;
  KfPtr fr1 = keyframe({feat_map_pyramid_1, feat_map_grad_pyramid_1, {}, {}, valid_mask_1, {}, {}}, camera_pyramid[0], L, FS, 8);
  const at::Tensor dp = sampled_dpts_0.to(at::kFloat).contiguous(), hm = sampled_locations_homo_0.to(at::kFloat).contiguous(),
                   sf = sampled_features_0.to(at::kFloat).contiguous();
  const auto R = hostf(rotation), t = hostf(translation), w = hostf(weights_tensor);
  std::vector<float> A(D * D), b(D);
  SAGE_OK(sage_ba_tracker_photo_jac_error(ctx(), fr1.get(), R.data(), t.data(), dp.data_ptr<float>(), hm.data_ptr<float>(),
                                          sf.data_ptr<float>(), (int)dp.numel(), with_scale, scale_0, eps, w.data(), A.data(), b.data(),
                                          &error, nullptr));
  AtA = to_dev(A.data(), {D, D}, rotation);
  Atb = to_dev(b.data(), {D, 1}, rotation);
}

template <int FS>
void tracker_photo_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation, const at::Tensor translation,
                                       const at::Tensor valid_mask_1, const at::Tensor sampled_dpts_0,
                                       const at::Tensor sampled_locations_homo_0, const at::Tensor sampled_features_0,
                                       const at::Tensor feat_map_pyramid_1, const at::Tensor feat_map_grad_pyramid_1,
                                       const at::Tensor level_offsets, const CameraPyramid<float> &camera_pyramid, const float eps,
                                       const at::Tensor weights_tensor)
{
  tracker_photo_jac<FS>(AtA, Atb, error, rotation, translation, valid_mask_1, sampled_dpts_0, sampled_locations_homo_0, sampled_features_0,
                        feat_map_pyramid_1, feat_map_grad_pyramid_1, level_offsets, camera_pyramid, 0, 0.f, eps, weights_tensor);
}

template <int FS>
void tracker_photo_jac_error_calculate_with_scale(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation,
                                                  const at::Tensor translation, const at::Tensor valid_mask_1,
                                                  const at::Tensor sampled_dpts_0, const at::Tensor sampled_locations_homo_0,
                                                  const at::Tensor sampled_features_0, const at::Tensor feat_map_pyramid_1,
                                                  const at::Tensor feat_map_grad_pyramid_1, const at::Tensor level_offsets,
                                                  const CameraPyramid<float> &camera_pyramid, const float scale_0, const float eps,
                                                  const at::Tensor weights_tensor)
{
  tracker_photo_jac<FS>(AtA, Atb, error, rotation, translation, valid_mask_1, sampled_dpts_0, sampled_locations_homo_0, sampled_features_0,
                        feat_map_pyramid_1, feat_map_grad_pyramid_1, level_offsets, camera_pyramid, 1, scale_0, eps, weights_tensor);
}

template <int FS>
float tracker_photo_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor valid_mask_1,
                                    const at::Tensor sampled_dpts_0, const at::Tensor sampled_locations_homo_0,
                                    const at::Tensor sampled_features_0, const at::Tensor feat_map_pyramid_1, const at::Tensor level_offsets,
                                    const CameraPyramid<float> &camera_pyramid, const float eps, const at::Tensor weights_tensor)
{
  const int L = (int)level_offsets.size(0);
  KfPtr fr1 = keyframe({feat_map_pyramid_1, {}, {}, {}, valid_mask_1, {}, {}}, camera_pyramid[0], L, FS, 8);
  const at::Tensor dp = sampled_dpts_0.to(at::kFloat).contiguous(), hm = sampled_locations_homo_0.to(at::kFloat).contiguous(),
                   sf = sampled_features_0.to(at::kFloat).contiguous();
  const auto R = hostf(rotation), t = hostf(translation), w = hostf(weights_tensor);
  float err = 0.f;
  SAGE_OK(sage_ba_tracker_photo_error(ctx(), fr1.get(), R.data(), t.data(), dp.data_ptr<float>(), hm.data_ptr<float>(), sf.data_ptr<float>(),
                                      (int)dp.numel(), eps, w.data(), &err, nullptr));
  return err;
}

// ------------------------------------------------------------------------------------------------ geometric
// The reference hands over KF1's depth map, its gradient and its basis as the caller computed them for the current
// code_1 / scale_1 (gtsam/geometric_factor.cpp:317-320,340-342).  The C ABI rebuilds the map and its central-difference
// gradient itself from (bias, basis, code, scale); feeding it bias := dpt_map_1 / scale_1, code := 0 reproduces the caller's
// tensors, so the call stays source-compatible.  (A caller that adopts the C ABI directly passes code_1 and drops those lines.)
template <int CS>
float geometric_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor flatten_dpt_map_bias_0,
                                const at::Tensor flatten_dpt_jac_code_0, const at::Tensor code_0, const at::Tensor dpt_map_1,
                                const at::Tensor valid_mask_1, const at::Tensor sampled_locations_1d_0,
                                const at::Tensor sampled_locations_homo_0, const float scale_0, const PinholeCamera<float> &camera,
                                const float eps, const float loss_param, const float weight)
{
  KfPtr kf0 = keyframe({{}, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, sampled_locations_1d_0, sampled_locations_homo_0}, camera,
                       1, 16, CS);
  const long HW = dpt_map_1.numel();
  static thread_local at::Tensor zero_basis; // the error-only operator needs no basis of KF1
  if (!zero_basis.defined() || zero_basis.size(0) != HW || zero_basis.device() != dpt_map_1.device())
    zero_basis = torch::zeros({HW, (long)CS}, dpt_map_1.options());
  const at::Tensor bias1 = dpt_map_1.to(at::kFloat).reshape({-1}).contiguous(), mask1 = valid_mask_1.to(at::kFloat).contiguous();
  KfPtr kf1 = borrowed_depth_keyframe(bias1, zero_basis, mask1, camera, CS);
  const auto R = hostf(rotation), t = hostf(translation), code = hostf(code_0);
  const std::vector<float> code1(CS, 0.f);
  float err = 0.f;
  SAGE_OK(sage_ba_geometric_error(ctx(), kf0.get(), kf1.get(), R.data(), t.data(), code.data(), code1.data(), scale_0, 1.0f, eps, loss_param,
                                  weight, &err, nullptr));
  return err;
}

template <int CS>
void geometric_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation10, const at::Tensor translation10,
                                   const at::Tensor rotation0, const at::Tensor translation0, const at::Tensor rotation1,
                                   const at::Tensor translation1, const at::Tensor flatten_dpt_map_bias_0,
                                   const at::Tensor flatten_dpt_jac_code_0, const at::Tensor code_0, const at::Tensor dpt_map_1,
                                   const at::Tensor /*dpt_map_grad_1: rebuilt inside by the same central differences*/,
                                   const at::Tensor dpt_jac_code_1, const at::Tensor valid_mask_1, const at::Tensor sampled_locations_1d_0,
                                   const at::Tensor sampled_locations_homo_0, const float scale_0, const float scale_1,
                                   const PinholeCamera<float> &camera, const float eps, const float loss_param, const float weight)
{
  KfPtr kf0 = keyframe({{}, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, sampled_locations_1d_0, sampled_locations_homo_0}, camera,
                       1, 16, CS);
  const at::Tensor unscaled = (dpt_map_1 / scale_1).to(at::kFloat).reshape({-1}).contiguous();
  const at::Tensor basis1 = dpt_jac_code_1.to(at::kFloat).reshape({-1, (long)CS}).contiguous(); // [H, W, C] -> [HW, C], as handed over
  const at::Tensor mask1 = valid_mask_1.to(at::kFloat).contiguous();
  KfPtr kf1 = borrowed_depth_keyframe(unscaled, basis1, mask1, camera, CS);
  constexpr int D = 14 + 2 * CS;
  std::vector<float> A(D * D), b(D);
  const HostPack h = hostpack({rotation10, translation10, rotation0, translation0, rotation1, translation1, code_0});
  const std::vector<float> code1(CS, 0.f);
  SAGE_OK(sage_ba_geometric_jac_error(ctx(), kf0.get(), kf1.get(), h[0], h[1], h[2], h[3], h[4], h[5], h[6], code1.data(), scale_0, scale_1,
                                      eps, loss_param, weight, A.data(), b.data(), &error, nullptr));
  AtA = to_dev(A.data(), {D, D}, rotation10);
  Atb = to_dev(b.data(), {D, 1}, rotation10);
}

// ------------------------------------------------------------------------------------------------ reprojection
void tracker_reproj_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation, const at::Tensor translation,
                                        const at::Tensor sampled_dpts_0, const at::Tensor sampled_locations_homo_0,
                                        const at::Tensor matched_locations_2d_1, const PinholeCamera<float> &camera, const float eps,
                                        const float loss_param, const float weight)
{
  const auto R = hostf(rotation), t = hostf(translation), dp = hostf(sampled_dpts_0), hm = hostf(sampled_locations_homo_0),
             uv = hostf(matched_locations_2d_1);
  const sage_ba_camera cam = cam_of(camera);
  std::vector<float> A(36), b(6);
  SAGE_OK(sage_ba_tracker_reproj_jac_error(ctx(), &cam, R.data(), t.data(), dp.data(), hm.data(), uv.data(), (int)dp.size(), eps, loss_param,
                                           weight, A.data(), b.data(), &error, nullptr));
  AtA = to_dev(A.data(), {6, 6}, rotation);
  Atb = to_dev(b.data(), {6, 1}, rotation);
}

float tracker_reproj_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor sampled_dpts_0,
                                     const at::Tensor sampled_locations_homo_0, const at::Tensor matched_locations_2d_1,
                                     const PinholeCamera<float> &camera, const float eps, const float loss_param, const float weight)
{
  const auto R = hostf(rotation), t = hostf(translation), dp = hostf(sampled_dpts_0), hm = hostf(sampled_locations_homo_0),
             uv = hostf(matched_locations_2d_1);
  const sage_ba_camera cam = cam_of(camera);
  float err = 0.f;
  SAGE_OK(sage_ba_tracker_reproj_error(ctx(), &cam, R.data(), t.data(), dp.data(), hm.data(), uv.data(), (int)dp.size(), eps, loss_param,
                                       weight, &err, nullptr));
  return err;
}

template <int CS>
void reprojection_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation10,
                                      const at::Tensor translation10, const at::Tensor rotation0, const at::Tensor translation0,
                                      const at::Tensor rotation1, const at::Tensor translation1, const at::Tensor flatten_dpt_map_bias_0,
                                      const at::Tensor flatten_dpt_jac_code_0, const at::Tensor code_0,
                                      const at::Tensor sampled_locations_1d_0, const at::Tensor sampled_locations_homo_0,
                                      const at::Tensor matched_locations_2d_1, const float scale_0, const PinholeCamera<float> &camera,
                                      const float eps, const float loss_param, const float weight)
{
  KfPtr kf0 = keyframe({{}, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, {}, {}}, camera, 1, 16, CS);
  constexpr int D = 13 + CS;
  std::vector<float> A(D * D), b(D);
  const auto R10 = hostf(rotation10), t10 = hostf(translation10), R0 = hostf(rotation0), t0 = hostf(translation0), R1 = hostf(rotation1),
             t1 = hostf(translation1), code = hostf(code_0), hm = hostf(sampled_locations_homo_0), uv = hostf(matched_locations_2d_1);
  const auto loc = hosti(sampled_locations_1d_0);
  SAGE_OK(sage_ba_reprojection_jac_error(ctx(), kf0.get(), R10.data(), t10.data(), R0.data(), t0.data(), R1.data(), t1.data(), code.data(),
                                         scale_0, loc.data(), hm.data(), uv.data(), (int)loc.size(), eps, loss_param, weight, A.data(),
                                         b.data(), &error, nullptr));
  AtA = to_dev(A.data(), {D, D}, rotation10);
  Atb = to_dev(b.data(), {D, 1}, rotation10);
}

template <int CS>
float reprojection_error_calculate(const at::Tensor rotation10, const at::Tensor translation10, const at::Tensor flatten_dpt_map_bias_0,
                                   const at::Tensor flatten_dpt_jac_code_0, const at::Tensor code_0, const at::Tensor sampled_locations_1d_0,
                                   const at::Tensor sampled_locations_homo_0, const at::Tensor matched_locations_2d_1, const float scale_0,
                                   const PinholeCamera<float> &camera, const float eps, const float loss_param, const float weight)
{
  KfPtr kf0 = keyframe({{}, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, {}, {}}, camera, 1, 16, CS);
  const auto R10 = hostf(rotation10), t10 = hostf(translation10), code = hostf(code_0), hm = hostf(sampled_locations_homo_0),
             uv = hostf(matched_locations_2d_1);
  const auto loc = hosti(sampled_locations_1d_0);
  float err = 0.f;
  SAGE_OK(sage_ba_reprojection_error(ctx(), kf0.get(), R10.data(), t10.data(), code.data(), scale_0, loc.data(), hm.data(), uv.data(),
                                     (int)loc.size(), eps, loss_param, weight, &err, nullptr));
  return err;
}

// ------------------------------------------------------------------------------------------------ match geometry
float tracker_match_geom_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor sampled_dpts_0,
                                         const at::Tensor matched_dpts_1, const at::Tensor sampled_locations_homo_0,
                                         const at::Tensor matched_locations_homo_1, const float loss_param, const float weight)
{
  const auto R = hostf(rotation), t = hostf(translation), d0 = hostf(sampled_dpts_0), d1 = hostf(matched_dpts_1),
             h0 = hostf(sampled_locations_homo_0), h1 = hostf(matched_locations_homo_1);
  float err = 0.f;
  SAGE_OK(sage_ba_tracker_match_geom_error(ctx(), R.data(), t.data(), d0.data(), d1.data(), h0.data(), h1.data(), (int)d0.size(), loss_param,
                                           weight, &err));
  return err;
}

static void tracker_mg_jac(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor &rotation, const at::Tensor &translation,
                           const at::Tensor &sampled_dpts_0, const at::Tensor &matched_dpts_1, const at::Tensor &sampled_locations_homo_0,
                           const at::Tensor &matched_locations_homo_1, int with_scale, float scale_0, float loss_param, float weight)
{
  const int D = with_scale ? 7 : 6;
  const auto R = hostf(rotation), t = hostf(translation), d0 = hostf(sampled_dpts_0), d1 = hostf(matched_dpts_1),
             h0 = hostf(sampled_locations_homo_0), h1 = hostf(matched_locations_homo_1);
  std::vector<float> A(D * D), b(D);
  SAGE_OK(sage_ba_tracker_match_geom_jac_error(ctx(), R.data(), t.data(), d0.data(), d1.data(), h0.data(), h1.data(), (int)d0.size(),
                                               with_scale, scale_0, loss_param, weight, A.data(), b.data(), &error));
  AtA = to_dev(A.data(), {D, D}, rotation);
  Atb = to_dev(b.data(), {D, 1}, rotation);
}

void tracker_match_geom_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation,
                                            const at::Tensor translation, const at::Tensor sampled_dpts_0, const at::Tensor matched_dpts_1,
                                            const at::Tensor sampled_locations_homo_0, const at::Tensor matched_locations_homo_1,
                                            const float loss_param, const float weight)
{
  tracker_mg_jac(AtA, Atb, error, rotation, translation, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0, matched_locations_homo_1, 0,
                 0.f, loss_param, weight);
}

void tracker_match_geom_jac_error_calculate_with_scale(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation,
                                                       const at::Tensor translation, const at::Tensor sampled_dpts_0,
                                                       const at::Tensor matched_dpts_1, const at::Tensor sampled_locations_homo_0,
                                                       const at::Tensor matched_locations_homo_1, const float scale_0,
                                                       const float loss_param, const float weight)
{
  tracker_mg_jac(AtA, Atb, error, rotation, translation, sampled_dpts_0, matched_dpts_1, sampled_locations_homo_0, matched_locations_homo_1, 1,
                 scale_0, loss_param, weight);
}

template <int CS>
float match_geometry_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor flatten_dpt_map_bias_0,
                                     const at::Tensor flatten_dpt_map_bias_1, const at::Tensor flatten_dpt_jac_code_0,
                                     const at::Tensor flatten_dpt_jac_code_1, const at::Tensor code_0, const at::Tensor code_1,
                                     const at::Tensor sampled_locations_homo_0, const at::Tensor matched_locations_homo_1,
                                     const at::Tensor sampled_locations_1d_0, const at::Tensor matched_locations_1d_1, const float scale_0,
                                     const float scale_1, const float loss_param, const float weight, const std::string robust_loss_type)
{
  const long HW = flatten_dpt_map_bias_0.numel();
  const PinholeCamera<float> cam(1.f, 1.f, 0.f, 0.f, (float)HW, 1.f); // only the pixel count matters for depth-only keyframes
  KfPtr kf0 = keyframe({{}, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, {}, {}}, cam, 1, 16, CS);
  KfPtr kf1 = keyframe({{}, {}, flatten_dpt_map_bias_1, flatten_dpt_jac_code_1, {}, {}, {}}, cam, 1, 16, CS);
  const auto R = hostf(rotation), t = hostf(translation), c0 = hostf(code_0), c1 = hostf(code_1), h0 = hostf(sampled_locations_homo_0),
             h1 = hostf(matched_locations_homo_1);
  const auto l0 = hosti(sampled_locations_1d_0), l1 = hosti(matched_locations_1d_1);
  float err = 0.f;
  SAGE_OK(sage_ba_match_geometry_error(ctx(), kf0.get(), kf1.get(), R.data(), t.data(), c0.data(), c1.data(), scale_0, scale_1, l0.data(),
                                       l1.data(), h0.data(), h1.data(), (int)l0.size(), loss_param, weight, loss_type(robust_loss_type), &err));
  return err;
}

template <int CS>
void match_geometry_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation10,
                                        const at::Tensor translation10, const at::Tensor rotation0, const at::Tensor translation0,
                                        const at::Tensor rotation1, const at::Tensor translation1, const at::Tensor flatten_dpt_map_bias_0,
                                        const at::Tensor flatten_dpt_map_bias_1, const at::Tensor flatten_dpt_jac_code_0,
                                        const at::Tensor flatten_dpt_jac_code_1, const at::Tensor code_0, const at::Tensor code_1,
                                        const at::Tensor sampled_locations_homo_0, const at::Tensor matched_locations_homo_1,
                                        const at::Tensor sampled_locations_1d_0, const at::Tensor matched_locations_1d_1, const float scale_0,
                                        const float scale_1, const float loss_param, const float weight, const std::string robust_loss_type)
{
  const long HW = flatten_dpt_map_bias_0.numel();
  const PinholeCamera<float> cam(1.f, 1.f, 0.f, 0.f, (float)HW, 1.f);
  KfPtr kf0 = keyframe({{}, {}, flatten_dpt_map_bias_0, flatten_dpt_jac_code_0, {}, {}, {}}, cam, 1, 16, CS);
  KfPtr kf1 = keyframe({{}, {}, flatten_dpt_map_bias_1, flatten_dpt_jac_code_1, {}, {}, {}}, cam, 1, 16, CS);
  constexpr int D = 14 + 2 * CS;
  std::vector<float> A(D * D), b(D);
  const auto R10 = hostf(rotation10), t10 = hostf(translation10), R0 = hostf(rotation0), t0 = hostf(translation0), R1 = hostf(rotation1),
             t1 = hostf(translation1), c0 = hostf(code_0), c1 = hostf(code_1), h0 = hostf(sampled_locations_homo_0),
             h1 = hostf(matched_locations_homo_1);
  const auto l0 = hosti(sampled_locations_1d_0), l1 = hosti(matched_locations_1d_1);
  SAGE_OK(sage_ba_match_geometry_jac_error(ctx(), kf0.get(), kf1.get(), R10.data(), t10.data(), R0.data(), t0.data(), R1.data(), t1.data(),
                                           c0.data(), c1.data(), scale_0, scale_1, l0.data(), l1.data(), h0.data(), h1.data(), (int)l0.size(),
                                           loss_param, weight, loss_type(robust_loss_type), A.data(), b.data(), &error));
  AtA = to_dev(A.data(), {D, D}, rotation10);
  Atb = to_dev(b.data(), {D, 1}, rotation10);
}

float loop_mg_error_calculate(const at::Tensor rotation, const at::Tensor translation, const at::Tensor sampled_unscaled_dpts_0,
                              const at::Tensor matched_unscaled_dpts_1, const at::Tensor sampled_locations_homo_0,
                              const at::Tensor matched_locations_homo_1, const float scale_0, const float scale_1, const float loss_param,
                              const float weight)
{
  const auto R = hostf(rotation), t = hostf(translation), d0 = hostf(sampled_unscaled_dpts_0), d1 = hostf(matched_unscaled_dpts_1),
             h0 = hostf(sampled_locations_homo_0), h1 = hostf(matched_locations_homo_1);
  float err = 0.f;
  SAGE_OK(sage_ba_loop_mg_error(ctx(), R.data(), t.data(), d0.data(), d1.data(), h0.data(), h1.data(), (int)d0.size(), scale_0, scale_1,
                                loss_param, weight, &err));
  return err;
}

void loop_mg_jac_error_calculate(at::Tensor &AtA, at::Tensor &Atb, float &error, const at::Tensor rotation10, const at::Tensor translation10,
                                 const at::Tensor rotation0, const at::Tensor translation0, const at::Tensor rotation1,
                                 const at::Tensor translation1, const at::Tensor sampled_unscaled_dpts_0,
                                 const at::Tensor matched_unscaled_dpts_1, const at::Tensor sampled_locations_homo_0,
                                 const at::Tensor matched_locations_homo_1, const float scale_0, const float scale_1, const float loss_param,
                                 const float weight)
{
  const auto R10 = hostf(rotation10), t10 = hostf(translation10), R0 = hostf(rotation0), t0 = hostf(translation0), R1 = hostf(rotation1),
             t1 = hostf(translation1), d0 = hostf(sampled_unscaled_dpts_0), d1 = hostf(matched_unscaled_dpts_1),
             h0 = hostf(sampled_locations_homo_0), h1 = hostf(matched_locations_homo_1);
  std::vector<float> A(14 * 14), b(14);
  SAGE_OK(sage_ba_loop_mg_jac_error(ctx(), R10.data(), t10.data(), R0.data(), t0.data(), R1.data(), t1.data(), d0.data(), d1.data(), h0.data(),
                                    h1.data(), (int)d0.size(), scale_0, scale_1, loss_param, weight, A.data(), b.data(), &error));
  AtA = to_dev(A.data(), {14, 14}, rotation10);
  Atb = to_dev(b.data(), {14, 1}, rotation10);
}

// explicit instantiations, as the reference's translation units end (cuda/photometric_factor_kernels.cpp:1388-1449 etc.)
template float photometric_error_calculate<DF_FEAT_SIZE>(const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                         const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                         const at::Tensor, const float, const CameraPyramid<float> &, const float,
                                                         const at::Tensor);
template void photometric_jac_error_calculate<DF_CODE_SIZE, DF_FEAT_SIZE>(at::Tensor &, at::Tensor &, float &, const at::Tensor, const at::Tensor,
                                                                          const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                                          const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                                          const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                                          const at::Tensor, const at::Tensor, const float,
                                                                          const CameraPyramid<float> &, const float, const at::Tensor);
template void tracker_photo_jac_error_calculate<DF_FEAT_SIZE>(at::Tensor &, at::Tensor &, float &, const at::Tensor, const at::Tensor,
                                                              const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                              const at::Tensor, const at::Tensor, const at::Tensor, const CameraPyramid<float> &,
                                                              const float, const at::Tensor);
template void tracker_photo_jac_error_calculate_with_scale<DF_FEAT_SIZE>(at::Tensor &, at::Tensor &, float &, const at::Tensor, const at::Tensor,
                                                                         const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                                         const at::Tensor, const at::Tensor, const at::Tensor,
                                                                         const CameraPyramid<float> &, const float, const float,
                                                                         const at::Tensor);
template float tracker_photo_error_calculate<DF_FEAT_SIZE>(const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                           const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                           const CameraPyramid<float> &, const float, const at::Tensor);
template float geometric_error_calculate<DF_CODE_SIZE>(const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                       const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const float,
                                                       const PinholeCamera<float> &, const float, const float, const float);
template void geometric_jac_error_calculate<DF_CODE_SIZE>(at::Tensor &, at::Tensor &, float &, const at::Tensor, const at::Tensor, const at::Tensor,
                                                          const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                          const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                          const at::Tensor, const at::Tensor, const float, const float,
                                                          const PinholeCamera<float> &, const float, const float, const float);
template void reprojection_jac_error_calculate<DF_CODE_SIZE>(at::Tensor &, at::Tensor &, float &, const at::Tensor, const at::Tensor,
                                                             const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                             const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                             const at::Tensor, const at::Tensor, const float, const PinholeCamera<float> &,
                                                             const float, const float, const float);
template float reprojection_error_calculate<DF_CODE_SIZE>(const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                          const at::Tensor, const at::Tensor, const at::Tensor, const float,
                                                          const PinholeCamera<float> &, const float, const float, const float);
template float match_geometry_error_calculate<DF_CODE_SIZE>(const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                            const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                            const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const float,
                                                            const float, const float, const float, const std::string);
template void match_geometry_jac_error_calculate<DF_CODE_SIZE>(at::Tensor &, at::Tensor &, float &, const at::Tensor, const at::Tensor,
                                                               const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                               const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                               const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
                                                               const at::Tensor, const at::Tensor, const float, const float, const float,
                                                               const float, const std::string);

} // namespace df

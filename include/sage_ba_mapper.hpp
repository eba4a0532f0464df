// sage_ba_mapper.hpp -- header-only C++ adapter that puts the batched LM of libsage_ba.so where Mapper::MappingStep /
// Mapper::UpdateMap stand in the reference, written against the reference's OWN map types.
//
//   sage::BatchedLocalBA<df::Map<float>, DF_CODE_SIZE> ba(map, opts);
//   ba.InitOneFrame(kf0->id);                       // Mapper::InitOneFrame's graph side   core/mapping/mapper.cpp:150-198
//   ba.EnqueueKeyframe(kf->id, conns);              // Mapper::EnqueueKeyframe             :300-380
//   ba.EnqueueLink(id0, id1, true, true, true);     // Mapper::EnqueueLink                 :395-445
//   sage_ba_lm_report r = ba.MappingStep();         // Mapper::MappingStep + UpdateMap     :469-612, :1141-1180
//
// MapT is duck-typed on exactly the members of df::Map / df::Keyframe / df::Frame the reference's mapper touches
// (core/mapping/keyframe_map.h:93-120, keyframe.h:19-61, frame.h:16-125): map->keyframes.Get(id); kf->id, pose_wk (Sophus::SE3),
// camera_pyramid_ptr, video_mask_ptr, feat_desc, feat_map_pyramid, feat_map_grad_pyramid, dpt_map_bias, dpt_jac_code (the strided
// [HW, C] view), code, dpt_scale, dpt_map, avg_squared_dpt_bias, sampled_locations_1d / _homo, valid_locations_1d,
// reinitialize_count, mutex.  Tensors are at::Tensor on the GPU, as the reference keeps them; nothing is copied through the host
// except the few hundred bytes of state.  oracle/build_ref.py compiles this header against the reference's keyframe_map.h
// (check_mapper_header.cpp) so that a drift of either side breaks the build, not the integration.
//
// What is different from the reference, by design: the solve is the device-side batched LM (every factor of the window is
// re-linearised each iteration; no ISAM2 incremental elimination, work items or marginalisation), and the descriptor matches of
// a reprojection factor are the cycle-consistent ones (sage_ba_cycle_match) without the TEASER++ filtering that follows in
// core/gtsam/reprojection_factor.cpp:136-186 (third-party, out of scope).
#pragma once
#include <torch/torch.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "sage_ba.h"

namespace sage
{

// the DeepFactorsOptions fields MappingStep's factors read (core/deepfactors_options.h, configs/slam_run.flags:96-106)
struct MapperOptions
{
  bool use_photometric = true, use_reprojection = true, use_geometric = true;
  std::vector<float> photo_factor_weights = {10.f, 9.f, 8.f, 7.f};
  float geo_factor_weight = 0.1f, geo_loss_param_factor = 0.03f;
  float reproj_factor_weight = 0.1f, reproj_loss_param_factor = 0.03f;
  float code_factor_weight = 1e-3f, init_scale_prior_weight = 1e-2f;
  float dpt_eps = 1e-4f;
  int desc_num_keypoints = 512;
  float desc_cyc_consis_thresh = 2.0f;
  int factor_iters = 10;
  double init_damp = 1e-4, min_damp = 1e-6, max_damp = 1e2, damp_dec_factor = 10.0, damp_inc_factor = 10.0;
};

template <typename MapT, int CS>
class BatchedLocalBA
{
public:
  using KeyframePtr = typename MapT::KeyframePtr;
  using FrameId = typename MapT::FrameId;

  // stream: the cudaStream_t the caller's tensors are produced on (nullptr: a private stream, synchronised per call).
  // rank / world / comm: multi-GPU sharding by keyframe owner (sage_ba_problem_set_shard); comm from sage_ba_comm_create.
  BatchedLocalBA(std::shared_ptr<MapT> map, const MapperOptions &opts, int device = 0, void *stream = nullptr, int rank = 0,
                 int world = 1, sage_ba_comm *comm = nullptr)
      : map_(std::move(map)), opts_(opts), rank_(rank), world_(world), comm_(comm)
  {
    if (sage_ba_create(&ctx_, device, stream) != 0)
      throw std::runtime_error("sage_ba_create failed: no usable CUDA device");
  }
  BatchedLocalBA(const BatchedLocalBA &) = delete;
  BatchedLocalBA &operator=(const BatchedLocalBA &) = delete;
  ~BatchedLocalBA()
  {
    for (auto &kv : dev_)
      sage_ba_keyframe_destroy(ctx_, kv.second);
    sage_ba_destroy(ctx_);
  }

  // first keyframe = the gauge: pose held, scale prior at its current scale, code prior (mapper.cpp:186-198)
  void InitOneFrame(FrameId id)
  {
    AddKeyframe(id);
    fixed_.push_back(id);
    scale_priors_.emplace_back(id, (float)map_->keyframes.Get(id)->dpt_scale);
    code_priors_.push_back(id);
  }

  // code prior + the enabled factor kinds in both directions for every back-connection (mapper.cpp:300-380)
  void EnqueueKeyframe(FrameId id, const std::vector<FrameId> &conns)
  {
    AddKeyframe(id);
    code_priors_.push_back(id);
    for (FrameId back : conns)
      EnqueueLink(id, back, true, true, true);
  }

  void EnqueueLink(FrameId id0, FrameId id1, bool photo, bool rep, bool geo)
  {
    const FrameId pair[2][2] = {{id0, id1}, {id1, id0}};
    for (const auto &ij : pair)
    {
      if (opts_.use_photometric && photo)
        factors_.push_back(Factor{0, ij[0], ij[1], {}, {}, {}, 0.f});
      if (opts_.use_reprojection && rep)
      {
        Factor f{2, ij[0], ij[1], {}, {}, {}, 0.f};
        if (CycleMatches(ij[0], ij[1], f))
          factors_.push_back(std::move(f));
      }
      if (opts_.use_geometric && geo) // loss scale from the link's first keyframe (mapper.cpp:357-360)
        factors_.push_back(Factor{1, ij[0], ij[1], {}, {}, {}, opts_.geo_loss_param_factor * (float)map_->keyframes.Get(id0)->avg_squared_dpt_bias});
    }
  }

  // optimise every variable the enqueued factors touch, then write the estimate back into the map
  sage_ba_lm_report MappingStep(int iters = 0)
  {
    const int K = (int)order_.size();
    std::map<FrameId, int> pos;
    std::vector<sage_ba_keyframe *> kfs(K);
    for (int k = 0; k < K; ++k)
    {
      pos[order_[k]] = k;
      kfs[k] = dev_.at(order_[k]);
    }
    sage_ba_problem *p = nullptr;
    Check(sage_ba_problem_create(ctx_, K, kfs.data(), &p));
    struct Guard
    {
      sage_ba_problem *p;
      ~Guard() { sage_ba_problem_destroy(p); }
    } guard{p};
    Check(sage_ba_problem_set_shard(p, rank_, world_));
    if (comm_)
      Check(sage_ba_problem_set_comm(p, comm_));
    KeyframePtr first = map_->keyframes.Get(order_[0]);
    const int L = (int)first->camera_pyramid_ptr->Levels();
    const float W = (float)(*first->camera_pyramid_ptr)[0].width();
    std::vector<float> w(opts_.photo_factor_weights.begin(), opts_.photo_factor_weights.begin() + L);
    for (const Factor &f : factors_)
    {
      if (f.kind == 0)
        Check(sage_ba_problem_add_photometric(p, pos.at(f.i), pos.at(f.j), w.data()));
      else if (f.kind == 1)
        Check(sage_ba_problem_add_geometric(p, pos.at(f.i), pos.at(f.j), f.loss, opts_.geo_factor_weight));
      else
        Check(sage_ba_problem_add_reprojection(p, pos.at(f.i), pos.at(f.j), f.loc.data(), f.homo.data(), f.uv.data(), (int)f.loc.size(),
                                               opts_.reproj_loss_param_factor * W * W, opts_.reproj_factor_weight));
    }
    for (FrameId id : code_priors_)
      Check(sage_ba_problem_add_code_prior(p, pos.at(id), nullptr, opts_.code_factor_weight));
    for (const auto &sp : scale_priors_)
      Check(sage_ba_problem_add_scale_prior(p, pos.at(sp.first), sp.second, opts_.init_scale_prior_weight));
    for (FrameId id : fixed_)
      Check(sage_ba_problem_fix(p, pos.at(id), 1, 0));
    // state in: pose_wk (Sophus::SE3 -> R row-major | t), code, dpt_scale
    std::vector<float> poses((size_t)K * 12), codes((size_t)K * CS), scales(K);
    for (int k = 0; k < K; ++k)
    {
      KeyframePtr kf = map_->keyframes.Get(order_[k]);
      std::shared_lock<std::shared_mutex> lock(kf->mutex);
      const auto R = kf->pose_wk.rotationMatrix();
      const auto t = kf->pose_wk.translation();
      for (int r = 0; r < 3; ++r)
      {
        for (int c = 0; c < 3; ++c)
          poses[(size_t)k * 12 + r * 3 + c] = (float)R(r, c);
        poses[(size_t)k * 12 + 9 + r] = (float)t(r);
      }
      const at::Tensor code = kf->code.to(at::kCPU, at::kFloat).reshape({-1}).contiguous();
      std::copy(code.data_ptr<float>(), code.data_ptr<float>() + CS, codes.begin() + (size_t)k * CS);
      scales[k] = (float)kf->dpt_scale;
    }
    Check(sage_ba_problem_set_state(p, poses.data(), codes.data(), scales.data(), opts_.dpt_eps));
    sage_ba_lm_options o{};
    o.max_iters = iters > 0 ? iters : opts_.factor_iters;
    o.init_damp = opts_.init_damp;
    o.min_damp = opts_.min_damp;
    o.max_damp = opts_.max_damp;
    o.damp_dec_factor = opts_.damp_dec_factor;
    o.damp_inc_factor = opts_.damp_inc_factor;
    o.min_rel_decrease = 1e-6;
    o.max_trials = 8;
    sage_ba_lm_report rep{};
    Check(sage_ba_problem_lm(p, &o, &rep));
    UpdateMap(p, K);
    return rep;
  }

private:
  struct Factor
  {
    int kind; // 0 photometric, 1 geometric, 2 reprojection
    FrameId i, j;
    std::vector<int32_t> loc;
    std::vector<float> homo, uv;
    float loss;
  };

  void Check(int rc) const
  {
    if (rc != 0)
      throw std::runtime_error(std::string("sage_ba: ") + sage_ba_last_error(ctx_));
  }

  // hand a keyframe's tensors to the library once (reference layouts in, re-laid-out on the device)
  void AddKeyframe(FrameId id)
  {
    if (dev_.count(id))
      return;
    KeyframePtr kf = map_->keyframes.Get(id);
    const auto &cam = (*kf->camera_pyramid_ptr)[0];
    sage_ba_keyframe_desc d{};
    d.memory = SAGE_BA_DEVICE;
    d.height = (int)cam.height();
    d.width = (int)cam.width();
    d.levels = (int)kf->camera_pyramid_ptr->Levels();
    d.feat_channels = (int)kf->feat_map_pyramid.size(0);
    d.code_size = CS;
    d.camera = {(float)cam.fx(), (float)cam.fy(), (float)cam.u0(), (float)cam.v0(), (float)cam.width(), (float)cam.height()};
    const at::Tensor feat = kf->feat_map_pyramid.to(at::kFloat).contiguous(), grad = kf->feat_map_grad_pyramid.to(at::kFloat).contiguous(),
                     bias = kf->dpt_map_bias.to(at::kFloat).reshape({-1}).contiguous(), mask = kf->video_mask_ptr->to(at::kFloat).contiguous(),
                     homo = kf->sampled_locations_homo.to(at::kFloat).contiguous(), loc = kf->sampled_locations_1d.to(at::kLong).contiguous();
    d.feat_map_pyramid = feat.data_ptr<float>();
    d.feat_map_grad_pyramid = grad.data_ptr<float>();
    d.dpt_map_bias = bias.data_ptr<float>();
    const at::Tensor jac = kf->dpt_jac_code; // the [HW, C] view with strides (1, HW) of the net's [C, H, W] output, used as is
    d.dpt_jac_code = jac.data_ptr<float>();
    d.jac_stride_row = jac.stride(0);
    d.jac_stride_col = jac.stride(1);
    d.video_mask = mask.data_ptr<float>();
    d.sampled_locations_1d = loc.data_ptr<int64_t>();
    d.sampled_locations_homo = homo.data_ptr<float>();
    d.num_samples = (int)loc.size(0);
    sage_ba_keyframe *h = nullptr;
    Check(sage_ba_keyframe_create(ctx_, &d, &h));
    dev_[id] = h;
    order_.push_back(id);
  }

  // ReprojectionFactor's constructor up to the TEASER step (core/gtsam/reprojection_factor.cpp:36-112): draw the keypoints with
  // std::shuffle(iota, mt19937(kf.id * fr.id)), cycle-match their descriptors on the device, keep the consistent ones
  bool CycleMatches(FrameId i, FrameId j, Factor &f)
  {
    KeyframePtr kf = map_->keyframes.Get(i), fr = map_->keyframes.Get(j);
    if (!kf->feat_desc.defined() || !fr->feat_desc.defined())
      return false;
    const at::Tensor valid = kf->valid_locations_1d.to(at::kCPU, at::kLong).contiguous();
    const long n = valid.size(0), K = std::min<long>(opts_.desc_num_keypoints, n);
    if (K <= 0)
      return false;
    std::vector<long> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::mt19937 g;
    g.seed((unsigned long)(kf->id * fr->id));
    std::shuffle(idx.begin(), idx.end(), g);
    std::vector<int64_t> kp(K);
    for (long k = 0; k < K; ++k)
      kp[k] = valid.data_ptr<int64_t>()[idx[k]];
    const at::Tensor d0 = kf->feat_desc.to(at::kFloat).contiguous(), d1 = fr->feat_desc.to(at::kFloat).contiguous();
    const auto &cam = (*fr->camera_pyramid_ptr)[0];
    const int H = (int)cam.height(), W = (int)cam.width(), Cd = (int)(d0.numel() / ((long)H * W));
    std::vector<int32_t> raw(K), cyc(K), inl(K);
    int ninl = 0;
    Check(sage_ba_cycle_match(ctx_, SAGE_BA_DEVICE, d0.data_ptr<float>(), d1.data_ptr<float>(), Cd, H, W, kp.data(), (int)K,
                              opts_.desc_cyc_consis_thresh, raw.data(), cyc.data(), inl.data(), &ninl, nullptr));
    if (ninl <= 0)
      return false;
    const auto &cam0 = (*kf->camera_pyramid_ptr)[0];
    for (int m = 0; m < ninl; ++m)
    {
      const int64_t l0 = kp[inl[m]];
      const int32_t l1 = raw[inl[m]];
      f.loc.push_back((int32_t)l0);
      const float u0 = (float)(l0 % W), v0 = (float)(l0 / W);
      f.homo.push_back((u0 - (float)cam0.u0()) / (float)cam0.fx()); // GenerateValidLocations' rays (mapping_utils.h:254-287)
      f.homo.push_back((v0 - (float)cam0.v0()) / (float)cam0.fy());
      f.homo.push_back(1.f);
      f.uv.push_back((float)(l1 % W));
      f.uv.push_back((float)(l1 / W));
    }
    return true;
  }

  // Mapper::UpdateMap (mapper.cpp:1141-1180): code, pose_wk, dpt_scale and UpdateDepth(...) -> dpt_map, under the keyframe's lock,
  // skipping keyframes that are being re-initialised
  void UpdateMap(sage_ba_problem *p, int K)
  {
    KeyframePtr first = map_->keyframes.Get(order_[0]);
    const auto &cam = (*first->camera_pyramid_ptr)[0];
    const long H = (long)cam.height(), W = (long)cam.width();
    std::vector<float> poses((size_t)K * 12), codes((size_t)K * CS), scales(K);
    at::Tensor maps = torch::empty({(long)K, H * W}, first->dpt_map_bias.options().dtype(at::kFloat));
    Check(sage_ba_problem_update_map(p, poses.data(), codes.data(), scales.data(), maps.data_ptr<float>(), SAGE_BA_DEVICE));
    for (int k = 0; k < K; ++k)
    {
      KeyframePtr kf = map_->keyframes.Get(order_[k]);
      std::unique_lock<std::shared_mutex> lock(kf->mutex);
      if (kf->reinitialize_count.load(std::memory_order_relaxed) > 0)
        continue;
      kf->code = torch::from_blob(codes.data() + (size_t)k * CS, {(long)CS, 1}, at::kFloat).clone().to(kf->dpt_map_bias.device());
      using SE3T = typename std::decay<decltype(kf->pose_wk)>::type;
      using Scalar = typename SE3T::Scalar;
      Eigen::Matrix<Scalar, 3, 3> R;
      Eigen::Matrix<Scalar, 3, 1> t;
      for (int r = 0; r < 3; ++r)
      {
        for (int c = 0; c < 3; ++c)
          R(r, c) = (Scalar)poses[(size_t)k * 12 + r * 3 + c];
        t(r) = (Scalar)poses[(size_t)k * 12 + 9 + r];
      }
      kf->pose_wk = SE3T(Eigen::Quaternion<Scalar>(R).normalized(), t);
      kf->dpt_scale = (Scalar)scales[k];
      kf->dpt_map = maps[k].reshape({H, W}).clone();
    }
  }

  std::shared_ptr<MapT> map_;
  MapperOptions opts_;
  int rank_, world_;
  sage_ba_comm *comm_;
  sage_ba_context *ctx_ = nullptr;
  std::map<FrameId, sage_ba_keyframe *> dev_;
  std::vector<FrameId> order_; // keyframe ids in insertion order = problem index
  std::vector<Factor> factors_;
  std::vector<FrameId> code_priors_, fixed_;
  std::vector<std::pair<FrameId, float>> scale_priors_;
};

} // namespace sage

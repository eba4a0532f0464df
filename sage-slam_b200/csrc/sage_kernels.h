// sage_kernels.h -- host-callable launchers of the factor kernels (internal; the public ABI is include/sage_ba.h).
#pragma once
#include <cuda_runtime.h>
#include "sage_common.cuh"

namespace sage
{
enum { PH_MAP_JAC = 0, PH_MAP_ERR = 1, PH_TRK_JAC = 2, PH_TRK_ERR = 3 };

// photometric.cu
int photo_row_width(int mode, int C);
int photo_samples_per_cta();
size_t photo_partial_floats(int mode, int C); // floats of one CTA's partial (lineariser modes: one WP x WP matrix per warp)
int photo_ctas_per_sm(int mode, int F, int C, bool staged = false);
int launch_photo(int mode, int F, int C, const PhotoFactor *factors, int nfactors, const CamPyr &cam, int slices, float *partH,
                 float *partE, float *out, int out_stride, int D, cudaStream_t stream, bool staged = false);

// geometric.cu
int geo_row_width(int C);
// tc: the caller's snapshot of geo_uses_tc(jac, C) -- partial size, slice count and launch must agree on the lineariser even if
// another thread flips the process-wide switch in between
bool geo_uses_tc(bool jac, int C);
int geo_set_tc(int on); // -1: query; returns the previous setting
int geo_ctas_per_sm(bool jac, int C, bool tc);
size_t geo_partial_floats(bool jac, int C, bool tc); // floats of one CTA's partial (the tcgen05 lineariser writes [128][112])
int launch_geo(bool jac, int C, const GeoFactor *factors, int nfactors, int W, int H, float fx, float fy, float cx, float cy,
               int slices, float *partH, float *partE, float *out, int out_stride, cudaStream_t stream, bool tc);

// reprojection.cu
int launch_reproj(bool jac, bool tracker, int C, const ReprojFactor *factors, int nfactors, float *out, int out_stride,
                  cudaStream_t stream);

int launch_match_geom(bool jac, const MatchGeomFactor *factors, int nfactors, float *out, int out_stride, int D, cudaStream_t stream);
int launch_map_match_geom(bool jac, int C, const MapMatchGeomFactor *factors, int nfactors, float *out, int out_stride, cudaStream_t stream);

// prep.cu
void launch_relayout_fg(const float *feat, const float *grad, float *fg, int F, long SP, cudaStream_t stream);
void launch_relayout_basis(const float *jac, long stride_row, long stride_col, float *basis, int HW, int C, cudaStream_t stream);
void launch_convert_loc(const int64_t *loc64, int *loc32, int N, int HW, int *bad, cudaStream_t stream);
void launch_permute_samples(const int *perm, const int *loc, const float4 *homo, int *loc_s, float4 *homo_s, int N, cudaStream_t stream);
void launch_pack_homo(const float *homo3, float4 *homo4, int N, cudaStream_t stream);
// dgm[HW] = (D, dD/dx, dD/dy, mask) with D = bias + basis . code (unscaled), central differences with replicate padding
void launch_depth_maps(const float *bias, const float *basis, const float *code_dev, const float *mask, float4 *dgm, float *scratch,
                       int H, int W, int C, cudaStream_t stream);
void launch_presample(const float *fg0, const float *bias0, const float *basis0, const int *loc1d, const float4 *homo,
                      const float *code_dev, float scale0, const CamPyr &cam, int F, int C, int N, float *out_dpts, float *out_homo,
                      float *out_feats, cudaStream_t stream);
int launch_build_pyramid(const float *feat, const float *mask0, float *fg, float *mask_scratch, const CamPyr &cam, int F, cudaStream_t stream);
} // namespace sage

// blocksolve.h -- block-sparse normal equations over keyframes: deterministic assembly, symbolic analysis and the
// left-looking block Cholesky that eliminates one keyframe's variables per CTA (blocksolve.cu).  Internal.
#pragma once
#include <cuda_runtime.h>

#include <utility>
#include <vector>

#include "sage_common.cuh"
#include "sage_internal.h"

namespace sage
{

struct FactorMeta
{
  int kind; // 0 photometric, 1 geometric, 2 reprojection
  int i, j;
  int D;
  int off;      // offset of [AtA | Atb | error | inliers] in the factor buffer
  int cost_off; // offset of [error | inliers] in the cost buffer
  int owner;    // rank that linearises the factor
};

struct PriorSpec
{
  int kind; // 0 code, 1 scale
  int kf;
  float weight;
  float init_scale;
  float init_code[SAGE_MAX_CODE];
};

// one assembled (non-fill) block of H: rows = variables of keyframe `runit`, columns = variables of `cunit`
struct AsmBlock
{
  int blk;
  int runit, cunit;
  int fbeg, fend; // contributing factors: asm_factors[fbeg, fend)
  int pbeg, pend; // priors on runit (diagonal blocks only): asm_priors[pbeg, pend)
};

// device view of the symbolic factorisation, passed by value to the kernels
struct BsDev
{
  int K;
  const int *unit_of_pos;  // [K]
  const int *col_ptr;      // [K+1] sub-diagonal blocks of column p: col_rowpos / col_blk [col_ptr[p], col_ptr[p+1])
  const int *col_rowpos;   // elimination position of the row keyframe, ascending
  const int *col_blk;      // block id
  const int *dep_ptr;      // [K+1] earlier columns j with L(p, j) != 0, ascending
  const int *dep_col;      // j
  const int *dep_blk;      // block id of L(p, j)
  const int *dep_pair_ptr; // [ndeps+1]
  const int *pair_src;     // block id of L(i, j), i >= p
  const int *pair_dst;     // slot of block (i, p) in column p: 0 = diagonal, s >= 1 = col_*[col_ptr[p] + s - 1]
  const unsigned char *fixed; // [K][SP] by keyframe; padding variables are marked fixed
};

struct BlockSystem
{
  int K = 0, C = 0, S = 0, SP = 0;
  int nblocks = 0;          // K diagonal + sub-diagonal blocks including fill
  int norig = 0;            // assembled blocks (diagonal + one per linked keyframe pair)
  int order_mode = 0;       // in: 1 forces the natural order; out: the candidate that won (0 natural, 1 forced, 2.. see build)
  double model_us = 0.0;    // modelled critical path of the factorisation with the chosen order
  int depth = 0;            // longest dependency chain of the elimination (critical path in block columns)
  long fill_blocks = 0;
  std::vector<int> order, pos;
  // host copies kept for diagnostics / the dense expansion
  std::vector<AsmBlock> asm_blocks_h;

  DevBuf<int> unit_of_pos, col_ptr, col_rowpos, col_blk, dep_ptr, dep_col, dep_blk, dep_pair_ptr, pair_src, pair_dst;
  DevBuf<int> asm_factors, asm_priors, sync;
  DevBuf<AsmBlock> asm_blocks;
  DevBuf<unsigned char> fixed;
  DevBuf<double> Hblk, Lblk, g, y, x, dinv;
  DevBuf<long long> dbg; // SAGE_BA_SOLVER_TRACE=1: per-column phase timestamps of the factorisation

  BsDev dev() const;
  // links: keyframe pairs that share a factor; fixed_vars: global variable order of include/sage_ba.h
  void build(int K_, int C_, const std::vector<FactorMeta> &metas, const std::vector<PriorSpec> &priors,
             const std::vector<unsigned char> &fixed_vars, int order_mode_, cudaStream_t s);
  // H blocks + g from the (complete) factor buffer and the priors at the given state, fixed summation order, fp64
  void assemble(const float *fbuf, const FactorMeta *metas_d, const PriorSpec *priors_d, const float *codes, const float *scales,
                cudaStream_t s, long *launches);
  // (H + damp diag H) delta = g with fixed variables held; delta in the global variable order.  info_d[0] != 0: not SPD.
  void solve(double damp, double *delta_d, int *info_d, cudaStream_t s, long *launches);
  // dense n x n copy of H (both triangles) and g in the global variable order (tests, the dense cross-check solver)
  void expand_dense(double *H, double *g, int n, cudaStream_t s, long *launches);
  int read_trace(long long *out, cudaStream_t s);
};

} // namespace sage

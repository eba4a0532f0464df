// blocksolve.cu -- the linear algebra of one LM iteration, hand-written and aware of the block sparsity of the problem.
//
// Stands where GTSAM's multifrontal Cholesky stands in the reference (core/mapping/mapper.cpp:544, isam_graph_->update):
// the normal equations of a keyframe graph are block-sparse with one (7 + C) x (7 + C) block per keyframe ([pose 6 | code C |
// scale]) on the diagonal and one per pair of keyframes that share a factor.  Eliminating keyframe k's block column IS the
// Schur complement of k's variables onto the keyframes it is linked to; doing it for every keyframe in an elimination order is
// a block Cholesky factorisation.  Three pieces:
//
//  * assemble_blocks_kernel: one CTA per non-zero block gathers the contributions of the factors touching that pair of
//    keyframes from the packed per-factor buffer, in the order the factors were added (fixed order, no atomics: the result is
//    bit-identical from run to run and for every GPU count) and adds the priors (core/gtsam/code_factor.cpp:42-104,
//    scale_factor.cpp:115-130).  Memory is O(K + links) blocks instead of the dense (K (7+C))^2 matrix.
//  * symbolic analysis (host, once per problem): elimination order = nested dissection over the keyframe index line
//    (temporal links make the graph banded: separators are b consecutive keyframes), fill blocks, and for every block
//    column the list of earlier columns it depends on.
//  * bs_factor_kernel: left-looking block Cholesky, ONE CTA PER BLOCK COLUMN, all columns in flight at once.  A column waits
//    (acquire on a per-column flag) only for the columns it really depends on, so independent subtrees of the elimination
//    tree run concurrently on different SMs and a dependent column overlaps its updates from older columns with the
//    factorisation of the newest one.  Inside a CTA: updates are 4x4 register-tiled fp64 block GEMMs on operands staged
//    k-major in shared memory, the diagonal block is factored by one warp holding rows in registers, the panel is solved one
//    row per thread.  The forward substitution rides along (the gradient is one more panel row); bs_backward_kernel walks the
//    tree in the opposite direction.  fp64 throughout: the damped system reaches condition numbers ~1e9.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>

#include "blocksolve.h"

namespace sage
{

template <int C>
struct BsCfg
{
  static constexpr int S = 7 + C;
  static constexpr int SP = (S + 7) / 8 * 8; // 40 / 24 / 16
  static constexpr int SPP = SP + 1;         // row stride of the panel blocks (odd: conflict-free column access)
  static constexpr int NT = 256;
  static constexpr int MAXS = C == 32 ? 7 : 16; // panel blocks resident in shared memory at a time
  static constexpr int TQ = SP / 4;
  static constexpr int TILES = TQ * TQ;
  static constexpr int NR = (SP + 31) / 32; // rows per lane in the warp-level factorisation
  static constexpr int BWS = C == 32 ? 12 : 16; // L(i, p) blocks of a column staged at once by the backward kernel
  static constexpr size_t factor_smem() { return sizeof(double) * ((size_t)MAXS * SP * SPP + 6 * SP * SP + 4 * SP); }
  static constexpr size_t backward_smem() { return sizeof(double) * ((size_t)SP * SPP + (size_t)BWS * SP * SP + (size_t)(BWS + 2) * SP); }
};

__device__ __forceinline__ int ld_acquire(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// 16-byte asynchronous global -> shared copies (LDGSTS, L2 only: the data was written by another SM in this launch)
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// one SP x SP fp64 block, all threads of the CTA
template <int BS, int NT>
__device__ __forceinline__ void block_copy_async(double *dst, const double *src, int tid)
{
  for (int e = tid * 2; e < BS; e += NT * 2)
    cp_async16(dst + e, src + e);
}

// column of factor `m` that holds variable `var` ([pose 6 | code C | scale]) of keyframe `unit`, or -1
__device__ __forceinline__ int factor_col(const FactorMeta &m, int unit, int var, int C)
{
  const bool is_i = unit == m.i;
  if (!is_i && unit != m.j)
    return -1;
  if (var < 6)
    return is_i ? var : 6 + var;
  const int c = var - 6;
  if (c < C)
    return is_i ? 12 + c : (m.kind == 1 ? 12 + C + c : -1);
  return is_i ? (m.kind == 1 ? 12 + 2 * C : 12 + C) : (m.kind == 1 ? 13 + 2 * C : -1);
}

__device__ __host__ __forceinline__ int global_var(int unit, int var, int K, int C)
{
  return var < 6 ? 6 * unit + var : 6 * K + unit * (C + 1) + (var - 6);
}

// ------------------------------------------------------------------------------------------------ assembly
__global__ void __launch_bounds__(256)
assemble_blocks_kernel(const float *__restrict__ fbuf, const FactorMeta *__restrict__ metas, const AsmBlock *__restrict__ blocks,
                       const int *__restrict__ asm_factors, const PriorSpec *__restrict__ priors, const int *__restrict__ asm_priors,
                       const float *__restrict__ codes, const float *__restrict__ scales, double *__restrict__ Hblk, double *__restrict__ g,
                       int C, int SP)
{
  const AsmBlock b = blocks[blockIdx.x];
  const int S = 7 + C;
  const bool diag = b.runit == b.cunit;
  // gridDim.y CTAs share the elements of a block (a diagonal block gathers from every factor of its keyframe)
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < SP * SP; e += blockDim.x * gridDim.y)
  {
    const int u = e / SP, v = e - u * SP;
    double sum = 0.0;
    if (u < S && v < S)
    {
      for (int q = b.fbeg; q < b.fend; ++q)
      {
        const FactorMeta m = metas[asm_factors[q]];
        const int lu = factor_col(m, b.runit, u, C), lv = factor_col(m, b.cunit, v, C);
        if (lu >= 0 && lv >= 0)
          sum += (double)fbuf[m.off + lu * m.D + lv];
      }
      if (diag && u == v && u >= 6)
        for (int q = b.pbeg; q < b.pend; ++q)
        {
          const PriorSpec &pr = priors[asm_priors[q]];
          if (pr.kind == 0 && u < 6 + C)
            sum += (double)pr.weight; // CodeFactor: AtA = w I
          else if (pr.kind == 1 && u == 6 + C)
          {
            const double s = (double)scales[pr.kf];
            sum += (double)pr.weight / (s * s); // ScaleFactor: AtA = w / s^2
          }
        }
    }
    Hblk[(size_t)b.blk * SP * SP + e] = sum;
  }
  if (diag && blockIdx.y == 0)
    for (int u = threadIdx.x; u < SP; u += blockDim.x)
    {
      double sum = 0.0;
      if (u < S)
      {
        for (int q = b.fbeg; q < b.fend; ++q)
        {
          const FactorMeta m = metas[asm_factors[q]];
          const int lu = factor_col(m, b.runit, u, C);
          if (lu >= 0)
            sum += (double)fbuf[m.off + m.D * m.D + lu];
        }
        if (u >= 6)
          for (int q = b.pbeg; q < b.pend; ++q)
          {
            const PriorSpec &pr = priors[asm_priors[q]];
            if (pr.kind == 0 && u < 6 + C)
              sum += (double)pr.weight * ((double)pr.init_code[u - 6] - (double)codes[pr.kf * C + (u - 6)]);
            else if (pr.kind == 1 && u == 6 + C)
            {
              const double s = (double)scales[pr.kf];
              sum += (double)pr.weight / s * (log((double)pr.init_scale) - log(s));
            }
          }
      }
      g[(size_t)b.runit * SP + u] = sum;
    }
}

__global__ void expand_dense_kernel(const AsmBlock *__restrict__ blocks, const double *__restrict__ Hblk, const double *__restrict__ gb,
                                    double *__restrict__ H, double *__restrict__ g, int n, int K, int C, int SP)
{
  const AsmBlock b = blocks[blockIdx.x];
  const int S = 7 + C;
  for (int e = threadIdx.x; e < S * S; e += blockDim.x)
  {
    const int u = e / S, v = e - u * S;
    const double val = Hblk[(size_t)b.blk * SP * SP + u * SP + v];
    const int gu = global_var(b.runit, u, K, C), gv = global_var(b.cunit, v, K, C);
    H[(size_t)gu * n + gv] = val;
    if (b.runit != b.cunit)
      H[(size_t)gv * n + gu] = val;
  }
  if (b.runit == b.cunit)
    for (int u = threadIdx.x; u < S; u += blockDim.x)
      g[global_var(b.runit, u, K, C)] = gb[(size_t)b.runit * SP + u];
}

// ------------------------------------------------------------------------------------------------ factorisation
// Compile-time recursion over the pivot column: every register-array index is a constant (a doubly nested `#pragma unroll` of
// 40 x 40 iterations is beyond what nvcc unrolls, and a dynamically indexed array would live in local memory).
template <int SP, int c>
struct TrsmStep
{
  __device__ __forceinline__ static void run(double (&a)[SP], const double *invd, const double *Lt)
  {
    const double xv = a[c] * invd[c];
    a[c] = xv;
#pragma unroll
    for (int qq = c + 1; qq < SP; ++qq)
      a[qq] = fma(-xv, Lt[c * SP + qq], a[qq]);
    TrsmStep<SP, c + 1>::run(a, invd, Lt);
  }
};
template <int SP>
struct TrsmStep<SP, SP>
{
  __device__ __forceinline__ static void run(double (&)[SP], const double *, const double *) {}
};

// Global storage of a factor block is TRANSPOSED (k-major): Lt[k * SP + r] = L[r][k], so that consumers stage it with a
// straight copy and read 4 consecutive rows of one column with two LDS.128.
template <int C>
__global__ void __launch_bounds__(BsCfg<C>::NT, 1)
bs_factor_kernel(const BsDev d, const double *__restrict__ Hblk, const double *__restrict__ g, double *__restrict__ Lblk,
                 double *__restrict__ ybuf, double *__restrict__ dinv, const double damp, int *__restrict__ sync,
                 long long *__restrict__ dbg)
{
  using T = BsCfg<C>;
  constexpr int SP = T::SP, SPP = T::SPP, NT = T::NT, MAXS = T::MAXS;
  // optional phase timestamps (globaltimer, ns) per column: [start, loaded, deps done, factored, panel solved, published, wait ns, deps]
  auto stamp = [&](int slot, int col) {
    if (dbg && threadIdx.x == 0)
    {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[(size_t)col * 8 + slot] = t;
    }
  };
  extern __shared__ __align__(16) double sm[];
  double *P = sm;                          // [MAXS][SP][SPP] panel blocks, row-major
  double *Bt = P + (size_t)MAXS * SP * SPP; // [SP][SP] k-major L(p, j)
  double *At = Bt + SP * SP;               // [2][2][SP][SP] k-major L(i, j): two pairs in use, two being prefetched
  double *Lt = At + 4 * SP * SP;           // [SP][SP] k-major L(p, p)
  double *gv = Lt + SP * SP;               // [SP] gradient / forward-substituted y_p
  double *yj = gv + SP;                    // [SP]
  double *invd = yj + SP;                  // [SP] 1 / L(p,p)[c][c]
  __shared__ int s_p;
  const int tid = threadIdx.x, lane = tid & 31;
  // columns are handed out in the order the CTAs actually start: a CTA only ever waits for columns that are already running
  if (tid == 0)
    s_p = atomicAdd(&sync[0], 1);
  __syncthreads();
  const int p = s_p;
  int *flags = sync + 2;
  const int unit = d.unit_of_pos[p];
  const int cbeg = d.col_ptr[p], nslots = d.col_ptr[p + 1] - cbeg + 1;
  const size_t BS = (size_t)SP * SP;
  stamp(0, p);
  long long waited = 0;
  // the column's dependency / pair lists, read many times below: keep them in shared memory when they fit
  constexpr int MAXDEP = 64, MAXPAIR = 512;
  __shared__ int s_dep_col[MAXDEP], s_dep_blk[MAXDEP], s_dep_pair[MAXDEP + 1], s_pair_src[MAXPAIR], s_pair_dst[MAXPAIR];
  const int dep0 = d.dep_ptr[p], ndep = d.dep_ptr[p + 1] - dep0;
  const int pair0 = d.dep_pair_ptr[dep0], npair = d.dep_pair_ptr[dep0 + ndep] - pair0;
  const bool meta_smem = ndep <= MAXDEP && npair <= MAXPAIR;
  if (meta_smem)
  {
    for (int e = tid; e < ndep; e += NT)
    {
      s_dep_col[e] = d.dep_col[dep0 + e];
      s_dep_blk[e] = d.dep_blk[dep0 + e];
    }
    for (int e = tid; e <= ndep; e += NT)
      s_dep_pair[e] = d.dep_pair_ptr[dep0 + e] - pair0;
    for (int e = tid; e < npair; e += NT)
    {
      s_pair_src[e] = d.pair_src[pair0 + e];
      s_pair_dst[e] = d.pair_dst[pair0 + e];
    }
  }
  const int *m_dep_col = meta_smem ? s_dep_col : d.dep_col + dep0;
  const int *m_dep_blk = meta_smem ? s_dep_blk : d.dep_blk + dep0;
  const int *m_pair_src = meta_smem ? s_pair_src : d.pair_src + pair0;
  const int *m_pair_dst = meta_smem ? s_pair_dst : d.pair_dst + pair0;

  for (int s0 = 0; s0 < nslots; s0 += MAXS)
  {
    const int s1 = min(nslots, s0 + MAXS);
    const bool first = s0 == 0;
    // ---- load this chunk of the column from H with the damping / gauge rules of the LM step applied
    for (int idx = tid; idx < (s1 - s0) * (int)BS; idx += NT)
    {
      const int sl = idx / (int)BS, e = idx - sl * (int)BS, u = e / SP, v = e - u * SP, slot = s0 + sl;
      const int blk = slot == 0 ? p : d.col_blk[cbeg + slot - 1];
      const int runit = slot == 0 ? unit : d.unit_of_pos[d.col_rowpos[cbeg + slot - 1]];
      double val = Hblk[(size_t)blk * BS + e];
      const bool fu = d.fixed[runit * SP + u], fv = d.fixed[unit * SP + v];
      if (slot == 0)
      {
        if (fu || fv)
          val = u == v ? 1.0 : 0.0; // fixed variable: identity row / column
        else if (u == v)
        {
          val = val + damp * val;
          if (!(val > 0.0))
            val = 1.0; // variable untouched by any factor: keep the system positive definite
        }
      }
      else if (fu || fv)
        val = 0.0;
      P[(size_t)sl * SP * SPP + u * SPP + v] = val;
    }
    if (first && tid < SP)
      gv[tid] = d.fixed[unit * SP + tid] ? 0.0 : g[(size_t)unit * SP + tid];
    __syncthreads();
    if (first)
      stamp(1, p);

    // ---- left-looking updates from every earlier column j with L(p, j) != 0.  Per dependency: wait for its flag, then ONE
    // round trip brings L(p, j), y_j and the first two L(i, j) (cp.async); further pairs are prefetched into the other buffer
    // pair while the current two are multiplied.
    for (int dep = 0; dep < ndep; ++dep)
    {
      const int j = m_dep_col[dep];
      if (first)
      {
        if (tid == 0)
        {
          long long t0 = 0, t1 = 0;
          if (dbg)
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          while (ld_acquire(&flags[j]) == 0)
            __nanosleep(20);
          if (dbg)
          {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            waited += t1 - t0;
          }
        }
        __syncthreads();
      }
      // pairs of this dependency whose destination lies in the resident chunk
      const int pb = meta_smem ? s_dep_pair[dep] : d.dep_pair_ptr[dep0 + dep] - pair0;
      const int pe = meta_smem ? s_dep_pair[dep + 1] : d.dep_pair_ptr[dep0 + dep + 1] - pair0;
      int q = pb;
      auto next_pair = [&]() -> int {
        while (q < pe)
        {
          const int dst = m_pair_dst[q];
          ++q;
          if (dst >= s0 && dst < s1)
            return q - 1;
        }
        return -1;
      };
      int cur[2] = {next_pair(), -1};
      if (!first && cur[0] < 0)
        continue; // nothing of this dependency lands in the resident chunk (uniform: every thread reads the same lists)
      block_copy_async<SP * SP, NT>(Bt, Lblk + (size_t)m_dep_blk[dep] * BS, tid);
      cur[1] = cur[0] >= 0 ? next_pair() : -1;
      int buf = 0;
      for (int h = 0; h < 2; ++h)
        if (cur[h] >= 0)
          block_copy_async<SP * SP, NT>(At + (size_t)(buf * 2 + h) * BS, Lblk + (size_t)m_pair_src[cur[h]] * BS, tid);
      cp_async_commit();
      if (first && tid < SP)
        yj[tid] = __ldcg(ybuf + (size_t)j * SP + tid);
      bool first_round = true;
      while (true)
      {
        int nxt[2] = {cur[0] >= 0 && cur[1] >= 0 ? next_pair() : -1, -1};
        nxt[1] = nxt[0] >= 0 ? next_pair() : -1;
        for (int h = 0; h < 2; ++h)
          if (nxt[h] >= 0)
            block_copy_async<SP * SP, NT>(At + (size_t)((buf ^ 1) * 2 + h) * BS, Lblk + (size_t)m_pair_src[nxt[h]] * BS, tid);
        cp_async_commit();
        cp_async_wait<1>(); // everything but the prefetch just issued
        __syncthreads();
        if (first_round && first && tid < SP)
        {
          double a = 0.0;
          for (int k = 0; k < SP; ++k)
            a = fma(Bt[k * SP + tid], yj[k], a);
          gv[tid] -= a; // g_p -= L(p, j) y_j
        }
        first_round = false;
        {
          // A(i, p) -= L(i, j) L(p, j)^T on the fp64 tensor cores: mma.sync m8n8k4, one 8x8 tile of the destination at a time.
          // Four warps per pair; a warp walks its tiles, SP / 4 DMMAs each.  Fragments straight from the k-major staging:
          //   a (row g, k t) = At[(k0 + t) * SP + r0 + g],  b (k t, col g) = Bt[(k0 + t) * SP + c0 + g],  c (row g, cols 2t, 2t+1)
          const int half = tid >> 7, w4 = (tid >> 5) & 3, g = lane >> 2, t4 = lane & 3;
          if (cur[half] >= 0)
          {
            constexpr int T8 = SP / 8;
            const double *A = At + (size_t)(buf * 2 + half) * BS;
            double *Cd = P + (size_t)(m_pair_dst[cur[half]] - s0) * SP * SPP;
            for (int tile = w4; tile < T8 * T8; tile += 4)
            {
              const int r0 = (tile / T8) * 8, c0 = (tile % T8) * 8;
              double d0 = 0.0, d1 = 0.0;
#pragma unroll
              for (int k0 = 0; k0 < SP; k0 += 4)
              {
                const double av = A[(k0 + t4) * SP + r0 + g], bv = Bt[(k0 + t4) * SP + c0 + g];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d0), "+d"(d1)
                             : "d"(av), "d"(bv));
              }
              double *c = Cd + (r0 + g) * SPP + c0 + 2 * t4;
              c[0] -= d0;
              c[1] -= d1;
            }
          }
        }
        __syncthreads(); // the buffers just read are overwritten by the next prefetch / the next dependency
        if (nxt[0] < 0)
          break;
        cur[0] = nxt[0];
        cur[1] = nxt[1];
        buf ^= 1;
      }
      cp_async_wait<0>();
    }

    // ---- factor the diagonal block with the whole CTA.  Square-root-free elimination with ONE barrier per pivot: the trailing
    // update a_rq -= a_rc a_qc / a_cc needs only the (unscaled) pivot column, which step c never writes, so scaling the columns
    // by 1 / sqrt(pivot) is deferred to one parallel pass at the end.  The serial chain per pivot is barrier + reciprocal + one
    // fused multiply-add (~250 cycles) instead of a warp walking the whole trailing block (2200 cycles measured).
    if (first)
    {
      stamp(2, p);
      // every thread keeps its share of the lower triangle (SP (SP + 1) / 2 elements, 4 per thread) in REGISTERS for the whole
      // elimination; per pivot the owners of column c publish it through a double-buffered shared-memory column, one barrier,
      // and everybody applies a_rq -= a_rc a_qc / a_cc to the elements it owns.
      constexpr int NE = SP * (SP + 1) / 2, EPT = (NE + NT - 1) / NT;
      double *D = P; // slot 0, row stride SPP
      double *colbuf = At; // [2][SP]: free until the next dependency round (none follows: the updates of this column are done)
      double *pivs = At + 2 * SP; // [SP]
      double av[EPT];
      int er[EPT], eq[EPT];
#pragma unroll
      for (int k = 0; k < EPT; ++k)
      {
        const int e = tid + k * NT;
        // e -> (r, q), q <= r: row r starts at r (r + 1) / 2
        int r = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) / 2 <= e)
          ++r;
        while (r * (r + 1) / 2 > e)
          --r;
        er[k] = e < NE ? r : -1;
        eq[k] = e < NE ? e - r * (r + 1) / 2 : -1;
        av[k] = e < NE ? D[r * SPP + eq[k]] : 0.0;
      }
      bool bad = false;
#pragma unroll 1
      for (int c = 0; c < SP; ++c)
      {
        double *cb = colbuf + (c & 1) * SP;
#pragma unroll
        for (int k = 0; k < EPT; ++k)
          if (eq[k] == c)
            cb[er[k]] = av[k];
        __syncthreads();
        double piv = cb[c];
        if (!(piv > 0.0))
        {
          bad = true;
          piv = 1.0;
        }
        if (tid == 0)
          pivs[c] = piv;
        const double rcp = __drcp_rn(piv);
#pragma unroll
        for (int k = 0; k < EPT; ++k)
          if (eq[k] > c)
            av[k] = fma(-cb[er[k]] * rcp, cb[eq[k]], av[k]);
      }
      if (bad && tid == 0)
        atomicMax(&sync[1], p + 1);
      __syncthreads();
      if (tid < SP)
        invd[tid] = rsqrt(pivs[tid]);
      for (int e = tid; e < SP * SP; e += NT)
        Lt[e] = 0.0;
      __syncthreads();
      // L[r][q] = a_rq / sqrt(pivot_q) (the diagonal: pivot / sqrt(pivot)), stored k-major
#pragma unroll
      for (int k = 0; k < EPT; ++k)
        if (er[k] >= 0)
          Lt[eq[k] * SP + er[k]] = (er[k] == eq[k] ? pivs[eq[k]] : av[k]) * invd[eq[k]];
      __syncthreads();
      stamp(3, p);
    }

    // ---- panel: X L(p,p)^T = A, one row per thread; the gradient is one more row (y_p = L(p,p)^-1 g_p)
    {
      const int sb = first ? 1 : s0;
      const int nrows = (s1 - sb) * SP + (first ? 1 : 0);
      for (int row = tid; row < nrows; row += NT)
      {
        const bool isg = first && row == nrows - 1;
        double *src = isg ? gv : P + (size_t)(sb - s0 + row / SP) * SP * SPP + (row % SP) * SPP;
        double a[SP];
#pragma unroll
        for (int c = 0; c < SP; ++c)
          a[c] = src[c];
        TrsmStep<SP, 0>::run(a, invd, Lt);
#pragma unroll
        for (int c = 0; c < SP; ++c)
          src[c] = a[c];
      }
    }
    __syncthreads();
    if (first)
      stamp(4, p);

    // ---- publish the chunk (k-major)
    for (int idx = tid; idx < (s1 - s0) * (int)BS; idx += NT)
    {
      const int sl = idx / (int)BS, e = idx - sl * (int)BS, k = e / SP, r = e - k * SP, slot = s0 + sl;
      const int blk = slot == 0 ? p : d.col_blk[cbeg + slot - 1];
      Lblk[(size_t)blk * BS + e] = slot == 0 ? Lt[e] : P[(size_t)sl * SP * SPP + r * SPP + k];
    }
    if (first && tid < SP)
    {
      ybuf[(size_t)p * SP + tid] = gv[tid];
      dinv[(size_t)p * SP + tid] = invd[tid];
    }
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (tid == 0)
    st_release(&flags[p], 1);
  stamp(5, p);
  if (dbg && tid == 0)
  {
    dbg[(size_t)p * 8 + 6] = waited;
    dbg[(size_t)p * 8 + 7] = ndep;
  }
}

// x_p = L(p,p)^-T (y_p - sum_{i in struct(p)} L(i,p)^T x_i), columns in reverse order, same flag protocol
template <int C>
__global__ void __launch_bounds__(BsCfg<C>::NT, 1)
bs_backward_kernel(const BsDev d, const double *__restrict__ Lblk, const double *__restrict__ ybuf, const double *__restrict__ dinv,
                   double *__restrict__ xbuf, double *__restrict__ delta, int Ccode, int *__restrict__ sync)
{
  using T = BsCfg<C>;
  constexpr int S = T::S, SP = T::SP, SPP = T::SPP, NT = T::NT, NR = T::NR;
  extern __shared__ __align__(16) double sm[];
  double *Lrow = sm;                             // [SP][SPP] L(p,p) row-major
  double *Bt = Lrow + SP * SPP;                  // [BWS][SP][SP] k-major L(i, p)
  double *acc = Bt + (size_t)T::BWS * SP * SP;   // [SP]
  double *xr = acc + SP;                         // [BWS][SP] x of the ancestors (then partial sums)
  __shared__ int s_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = d.K;
  int *ticket = sync + 2 + K, *flags = sync + 3 + K;
  if (tid == 0)
    s_p = K - 1 - atomicAdd(ticket, 1);
  __syncthreads();
  const int p = s_p;
  const int unit = d.unit_of_pos[p];
  const int cbeg = d.col_ptr[p], nr = d.col_ptr[p + 1] - cbeg;
  const size_t BS = (size_t)SP * SP;
  // the factor is final (the forward kernel ended): stage L(p,p) and the column's sub-diagonal blocks BEFORE looking at any flag
  constexpr int BWS = T::BWS;
  for (int c0 = 0; c0 < max(nr, 1); c0 += BWS)
  {
    const int c1 = min(nr, c0 + BWS);
    if (c0 == 0)
    {
      for (int e = tid; e < (int)BS; e += NT)
      {
        const int k = e / SP, r = e - k * SP;
        Lrow[r * SPP + k] = __ldcg(Lblk + (size_t)p * BS + e);
      }
      if (tid < SP)
        acc[tid] = __ldcg(ybuf + (size_t)p * SP + tid);
    }
    for (int s = c0; s < c1; ++s)
      block_copy_async<SP * SP, NT>(Bt + (size_t)(s - c0) * BS, Lblk + (size_t)d.col_blk[cbeg + s] * BS, tid);
    cp_async_commit();
    // x of the ancestors: wait for each flag (they complete top-down; usually all but the last are long done)
    for (int s = c0 + warp; s < c1; s += NT / 32)
    {
      const int rp = d.col_rowpos[cbeg + s];
      if (lane == 0)
        while (ld_acquire(&flags[rp]) == 0)
          __nanosleep(20);
      __syncwarp();
      for (int r = lane; r < SP; r += 32)
        xr[(size_t)(s - c0) * SP + r] = __ldcg(xbuf + (size_t)rp * SP + r);
    }
    cp_async_wait<0>();
    __syncthreads();
    // acc[k] -= sum_s sum_r L_s[r][k] x_s[r]: thread (k, part) sums its share of the slots, the parts are folded in a fixed order
    {
      constexpr int PARTS = NT / SP; // 6 for SP = 40
      const int k = tid % SP, part = tid / SP;
      double a = 0.0;
      if (part < PARTS)
        for (int s = c0 + part; s < c1; s += PARTS)
        {
          const double *B = Bt + (size_t)(s - c0) * BS + k * SP, *x = xr + (size_t)(s - c0) * SP;
          for (int r = 0; r < SP; ++r)
            a = fma(B[r], x[r], a);
        }
      __syncthreads(); // xr is reused as the partial-sum scratch: [PARTS][SP]
      if (part < PARTS)
        xr[part * SP + k] = a;
      __syncthreads();
      if (tid < SP)
      {
        double t = 0.0;
        for (int q = 0; q < PARTS; ++q)
          t += xr[q * SP + tid];
        acc[tid] -= t;
      }
      __syncthreads();
    }
  }
  if (warp == 0)
  {
    double a[NR], di[NR];
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
    {
      a[rr] = lane + 32 * rr < SP ? acc[lane + 32 * rr] : 0.0;
      di[rr] = lane + 32 * rr < SP ? __ldcg(dinv + (size_t)p * SP + lane + 32 * rr) : 0.0;
    }
#pragma unroll
    for (int c = SP - 1; c >= 0; --c)
    {
      const double xc = __shfl_sync(0xffffffffu, a[c >> 5] * di[c >> 5], c & 31);
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
      {
        const int qq = lane + 32 * rr;
        if (qq == c)
          a[rr] = xc;
        else if (qq < c)
          a[rr] = fma(-Lrow[c * SPP + qq], xc, a[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
    {
      const int r = lane + 32 * rr;
      if (r < SP)
      {
        xbuf[(size_t)p * SP + r] = a[rr];
        if (r < S)
          delta[global_var(unit, r, K, Ccode)] = a[rr];
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0)
    st_release(&flags[p], 1);
}

// ------------------------------------------------------------------------------------------------ host side
BsDev BlockSystem::dev() const
{
  BsDev d;
  d.K = K;
  d.unit_of_pos = unit_of_pos.p;
  d.col_ptr = col_ptr.p;
  d.col_rowpos = col_rowpos.p;
  d.col_blk = col_blk.p;
  d.dep_ptr = dep_ptr.p;
  d.dep_col = dep_col.p;
  d.dep_blk = dep_blk.p;
  d.dep_pair_ptr = dep_pair_ptr.p;
  d.pair_src = pair_src.p;
  d.pair_dst = pair_dst.p;
  d.fixed = fixed.p;
  return d;
}

static void nested_dissection(int lo, int hi, int b, std::vector<int> &out)
{
  const int n = hi - lo;
  if (n <= std::max(3 * b, 8))
  {
    for (int u = lo; u < hi; ++u)
      out.push_back(u);
    return;
  }
  const int s0 = lo + n / 2 - b / 2, s1 = s0 + b;
  nested_dissection(lo, s0, b, out);
  nested_dissection(s1, hi, b, out);
  for (int u = s0; u < s1; ++u)
    out.push_back(u);
}

template <typename V>
static void upload(DevBuf<typename V::value_type> &dst, const V &src, cudaStream_t s)
{
  using E = typename V::value_type;
  dst.ensure(std::max<size_t>(src.size(), 1));
  if (!src.empty())
    SAGE_CUDA(cudaMemcpyAsync((void *)dst.p, (const void *)src.data(), src.size() * sizeof(E), cudaMemcpyHostToDevice, s));
}

void BlockSystem::build(int K_, int C_, const std::vector<FactorMeta> &metas, const std::vector<PriorSpec> &priors,
                        const std::vector<unsigned char> &fixed_vars, int order_mode_, cudaStream_t s)
{
  K = K_;
  C = C_;
  S = 7 + C;
  SP = (S + 7) / 8 * 8;
  order_mode = order_mode_;
  // ---- keyframe graph
  std::set<std::pair<int, int>> links;
  std::vector<int> span;
  for (const FactorMeta &m : metas)
  {
    const int a = std::min(m.i, m.j), b = std::max(m.i, m.j);
    if (links.insert({a, b}).second)
      span.push_back(b - a);
  }
  // ---- elimination order: three candidates -- natural (temporal), nested dissection over the keyframe index line (band
  // graphs: separators are b consecutive keyframes, subtrees run concurrently) and minimum degree (graphs with long-range
  // links: far less fill) -- compared by a model of the critical path of the factorisation (per column: a fixed cost, the rows
  // it solves, and the update rounds it must wait for from its last dependency); the cheapest wins.
  auto symbolic = [&](const std::vector<int> &ord, std::vector<std::set<int>> &cs_out, double &crit) {
    std::vector<int> ps(K);
    for (int q = 0; q < K; ++q)
      ps[ord[q]] = q;
    cs_out.assign(K, {});
    for (const auto &l : links)
    {
      const int a = ps[l.first], b = ps[l.second];
      cs_out[std::min(a, b)].insert(std::max(a, b));
    }
    for (int q = 0; q < K; ++q)
      if (!cs_out[q].empty())
      {
        auto it = cs_out[q].begin();
        const int parent = *it;
        for (++it; it != cs_out[q].end(); ++it)
          cs_out[parent].insert(*it);
      }
    std::vector<std::vector<int>> dp(K);
    for (int j = 0; j < K; ++j)
      for (int r : cs_out[j])
        dp[r].push_back(j);
    std::vector<double> cost(K, 0.0);
    crit = 0.0;
    for (int q = 0; q < K; ++q)
    {
      double before = 0.0;
      int pairs_last = 0;
      for (int j : dp[q])
      {
        if (cost[j] > before)
        {
          before = cost[j];
          pairs_last = 0;
          for (int i : cs_out[j])
            pairs_last += i >= q ? 1 : 0;
        }
      }
      const int chunks = ((int)cs_out[q].size() + 7) / 7; // a column wider than the resident panel walks its dependencies again
      cost[q] = before + 16.0 + 0.5 * (double)cs_out[q].size() + 1.4 * ((pairs_last + 1) / 2) + 3.0 * (chunks - 1) * (double)dp[q].size();
      crit = std::max(crit, cost[q]);
    }
  };
  std::vector<std::vector<int>> candidates;
  {
    std::vector<int> nat(K);
    for (int u = 0; u < K; ++u)
      nat[u] = u;
    candidates.push_back(nat);
    if (order_mode != 1 && !span.empty())
    {
      std::vector<int> sp = span;
      std::sort(sp.begin(), sp.end());
      const int b = std::max(1, sp[std::min(sp.size() - 1, (size_t)(0.9 * sp.size()))]); // band of the temporal links
      std::vector<int> ndo;
      nested_dissection(0, K, b, ndo);
      candidates.push_back(ndo);
      // minimum degree on the keyframe graph (exact, with explicit fill; K is a few hundred at most)
      std::vector<std::set<int>> adj(K);
      for (const auto &l : links)
      {
        adj[l.first].insert(l.second);
        adj[l.second].insert(l.first);
      }
      std::vector<char> gone(K, 0);
      std::vector<int> md;
      for (int step = 0; step < K; ++step)
      {
        int best = -1;
        for (int u = 0; u < K; ++u)
          if (!gone[u] && (best < 0 || adj[u].size() < adj[best].size()))
            best = u;
        const std::vector<int> nb(adj[best].begin(), adj[best].end());
        for (int a : nb)
        {
          adj[a].erase(best);
          for (int b2 : nb)
            if (a != b2)
              adj[a].insert(b2);
        }
        gone[best] = 1;
        md.push_back(best);
      }
      candidates.push_back(md);
    }
  }
  std::vector<std::set<int>> cs;
  {
    double best = 0.0;
    int pick = 0;
    for (size_t c = 0; c < candidates.size(); ++c)
    {
      std::vector<std::set<int>> tmp;
      double crit = 0.0;
      symbolic(candidates[c], tmp, crit);
      if (c == 0 || crit < best)
      {
        best = crit;
        pick = (int)c;
      }
    }
    order = candidates[pick];
    order_mode = order_mode == 1 ? 1 : pick; // 0 natural won, 1 forced natural, 2 nested dissection, ... reported by solver_info
    double crit = 0.0;
    symbolic(order, cs, crit);
    model_us = crit;
  }
  pos.assign(K, 0);
  for (int q = 0; q < K; ++q)
    pos[order[q]] = q;
  std::set<std::pair<int, int>> orig; // (row pos, col pos) of assembled sub-diagonal blocks
  for (const auto &l : links)
  {
    const int a = pos[l.first], b = pos[l.second];
    orig.insert({std::max(a, b), std::min(a, b)});
  }
  std::vector<int> h_col_ptr(K + 1, 0), h_rowpos, h_blk;
  for (int q = 0; q < K; ++q)
  {
    for (int r : cs[q])
    {
      h_rowpos.push_back(r);
      h_blk.push_back(K + (int)h_blk.size());
    }
    h_col_ptr[q + 1] = (int)h_rowpos.size();
  }
  nblocks = K + (int)h_rowpos.size();
  fill_blocks = (long)h_rowpos.size() - (long)orig.size();
  auto block_id = [&](int r, int c) -> int {
    if (r == c)
      return c;
    for (int e = h_col_ptr[c]; e < h_col_ptr[c + 1]; ++e)
      if (h_rowpos[e] == r)
        return h_blk[e];
    return -1;
  };
  auto slot_of = [&](int r, int c) -> int {
    if (r == c)
      return 0;
    for (int e = h_col_ptr[c]; e < h_col_ptr[c + 1]; ++e)
      if (h_rowpos[e] == r)
        return e - h_col_ptr[c] + 1;
    return -1;
  };
  // ---- dependency lists (left-looking): column p needs every earlier column j with p in struct(j)
  std::vector<std::vector<int>> rowlist(K);
  for (int j = 0; j < K; ++j)
    for (int r : cs[j])
      rowlist[r].push_back(j);
  std::vector<int> h_dep_ptr(K + 1, 0), h_dep_col, h_dep_blk, h_dep_pair_ptr(1, 0), h_pair_src, h_pair_dst, chain(K, 1);
  depth = K ? 1 : 0;
  for (int q = 0; q < K; ++q)
  {
    // a column consumes its dependencies in a FIXED order (the order of the floating-point updates must not depend on timing);
    // sorting them by the length of their own dependency chain puts the ones that finish last at the end, so the column does not
    // sit behind a late dependency while ten finished ones are queued after it
    std::stable_sort(rowlist[q].begin(), rowlist[q].end(), [&](int a, int b) { return chain[a] < chain[b]; });
    for (int j : rowlist[q])
    {
      h_dep_col.push_back(j);
      h_dep_blk.push_back(block_id(q, j));
      for (int i : cs[j])
        if (i >= q)
        {
          const int sl = slot_of(i, q);
          SAGE_CHECK(sl >= 0, "symbolic factorisation: missing fill block");
          h_pair_src.push_back(block_id(i, j));
          h_pair_dst.push_back(sl);
        }
      h_dep_pair_ptr.push_back((int)h_pair_src.size());
      chain[q] = std::max(chain[q], chain[j] + 1);
    }
    h_dep_ptr[q + 1] = (int)h_dep_col.size();
    depth = std::max(depth, chain[q]);
  }
  // ---- assembly lists: diagonal blocks and one block per linked pair, contributors in the order the factors were added
  asm_blocks_h.clear();
  std::vector<int> h_asm_factors, h_asm_priors;
  for (int u = 0; u < K; ++u)
  {
    AsmBlock b;
    b.blk = pos[u];
    b.runit = b.cunit = u;
    b.fbeg = (int)h_asm_factors.size();
    for (size_t f = 0; f < metas.size(); ++f)
      if (metas[f].i == u || metas[f].j == u)
        h_asm_factors.push_back((int)f);
    b.fend = (int)h_asm_factors.size();
    b.pbeg = (int)h_asm_priors.size();
    for (size_t q = 0; q < priors.size(); ++q)
      if (priors[q].kf == u)
        h_asm_priors.push_back((int)q);
    b.pend = (int)h_asm_priors.size();
    asm_blocks_h.push_back(b);
  }
  for (const auto &l : links)
  {
    const int a = l.first, c = l.second;
    const bool a_row = pos[a] > pos[c];
    AsmBlock b;
    b.runit = a_row ? a : c;
    b.cunit = a_row ? c : a;
    b.blk = block_id(pos[b.runit], pos[b.cunit]);
    SAGE_CHECK(b.blk >= 0, "symbolic factorisation: missing link block");
    b.fbeg = (int)h_asm_factors.size();
    for (size_t f = 0; f < metas.size(); ++f)
      if ((metas[f].i == a && metas[f].j == c) || (metas[f].i == c && metas[f].j == a))
        h_asm_factors.push_back((int)f);
    b.fend = (int)h_asm_factors.size();
    b.pbeg = b.pend = 0;
    asm_blocks_h.push_back(b);
  }
  norig = (int)asm_blocks_h.size();
  // ---- fixed variables per keyframe, padding marked fixed
  std::vector<unsigned char> h_fixed((size_t)K * SP, 1);
  for (int u = 0; u < K; ++u)
    for (int v = 0; v < S; ++v)
    {
      const int gi = global_var(u, v, K, C);
      h_fixed[(size_t)u * SP + v] = gi < (int)fixed_vars.size() ? fixed_vars[gi] : 0;
    }
  upload(unit_of_pos, order, s);
  upload(col_ptr, h_col_ptr, s);
  upload(col_rowpos, h_rowpos, s);
  upload(col_blk, h_blk, s);
  upload(dep_ptr, h_dep_ptr, s);
  upload(dep_col, h_dep_col, s);
  upload(dep_blk, h_dep_blk, s);
  upload(dep_pair_ptr, h_dep_pair_ptr, s);
  upload(pair_src, h_pair_src, s);
  upload(pair_dst, h_pair_dst, s);
  upload(asm_factors, h_asm_factors, s);
  upload(asm_priors, h_asm_priors, s);
  upload(asm_blocks, asm_blocks_h, s);
  upload(fixed, h_fixed, s);
  const size_t BS = (size_t)SP * SP;
  Hblk.ensure((size_t)nblocks * BS);
  Lblk.ensure((size_t)nblocks * BS);
  g.ensure((size_t)K * SP);
  y.ensure((size_t)K * SP);
  x.ensure((size_t)K * SP);
  dinv.ensure((size_t)K * SP);
  sync.ensure(4 + 2 * (size_t)K);
  SAGE_CUDA(cudaMemsetAsync(Hblk.p, 0, sizeof(double) * nblocks * BS, s)); // fill blocks stay zero
  SAGE_CUDA(cudaStreamSynchronize(s)); // the host vectors above go out of scope
}

void BlockSystem::assemble(const float *fbuf, const FactorMeta *metas_d, const PriorSpec *priors_d, const float *codes, const float *scales,
                           cudaStream_t s, long *launches)
{
  if (norig <= 0)
    return;
  assemble_blocks_kernel<<<dim3(norig, 6), 256, 0, s>>>(fbuf, metas_d, asm_blocks.p, asm_factors.p, priors_d, asm_priors.p, codes, scales, Hblk.p,
                                               g.p, C, SP);
  if (launches)
    *launches += 1;
}

template <int C>
static void solve_t(BlockSystem &b, double damp, double *delta_d, cudaStream_t s)
{
  using T = BsCfg<C>;
  static unsigned long long done = 0;
  if (first_use_on_device(done))
  {
    cudaFuncSetAttribute(bs_factor_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::factor_smem());
    cudaFuncSetAttribute(bs_backward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::backward_smem());
  }
  long long *dbg = nullptr;
  if (getenv("SAGE_BA_SOLVER_TRACE"))
    dbg = reinterpret_cast<long long *>(b.dbg.ensure((size_t)b.K * 8));
  bs_factor_kernel<C><<<b.K, T::NT, T::factor_smem(), s>>>(b.dev(), b.Hblk.p, b.g.p, b.Lblk.p, b.y.p, b.dinv.p, damp, b.sync.p, dbg);
  bs_backward_kernel<C><<<b.K, T::NT, T::backward_smem(), s>>>(b.dev(), b.Lblk.p, b.y.p, b.dinv.p, b.x.p, delta_d, C, b.sync.p);
}

void BlockSystem::solve(double damp, double *delta_d, int *info_d, cudaStream_t s, long *launches)
{
  SAGE_CUDA(cudaMemsetAsync(sync.p, 0, sizeof(int) * (4 + 2 * (size_t)K), s));
  switch (C)
  {
  case 32: solve_t<32>(*this, damp, delta_d, s); break;
  case 16: solve_t<16>(*this, damp, delta_d, s); break;
  case 8: solve_t<8>(*this, damp, delta_d, s); break;
  default: SAGE_CHECK(false, "unsupported code_size");
  }
  if (info_d)
    SAGE_CUDA(cudaMemcpyAsync(info_d, sync.p + 1, sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (launches)
    *launches += 2;
}

// phase timestamps of the last factorisation (ns since the first column started), row = elimination position:
// [start, loaded, deps done, factored, panel solved, published, ns spent waiting on flags, number of dependencies]
int BlockSystem::read_trace(long long *out, cudaStream_t s)
{
  if (!dbg.p)
    return 0;
  SAGE_CUDA(cudaMemcpyAsync(out, dbg.p, sizeof(long long) * K * 8, cudaMemcpyDeviceToHost, s));
  SAGE_CUDA(cudaStreamSynchronize(s));
  return K;
}

void BlockSystem::expand_dense(double *H, double *gd, int n, cudaStream_t s, long *launches)
{
  SAGE_CUDA(cudaMemsetAsync(H, 0, sizeof(double) * (size_t)n * n, s));
  SAGE_CUDA(cudaMemsetAsync(gd, 0, sizeof(double) * n, s));
  if (norig > 0)
    expand_dense_kernel<<<norig, 256, 0, s>>>(asm_blocks.p, Hblk.p, g.p, H, gd, n, K, C, SP);
  if (launches)
    *launches += 1;
}

} // namespace sage

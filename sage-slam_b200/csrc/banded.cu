// banded.cu -- block-banded Cholesky solve of the damped normal equations for chain-shaped covisibility.
//
// With temporal links only (core/mapping/mapper.cpp:339-375: a keyframe connects to its <= b predecessors) the
// Hessian is block-banded once the variables are interleaved per keyframe, x_k = [pose_k 6 | code_k C | scale_k]:
// block (i, j) is non-zero only for |i - j| <= b.  One CTA factorises the K x K block matrix (block size S = 7 + C)
// in fp64 with a right-looking block Cholesky and does the forward / backward substitution -- ~30 MFLOP for
// K = 32, b = 3, C = 32 instead of the dense 1056^3/3 of the generic Schur + cuSOLVER path (problem.cu), and a single
// launch instead of ~25 library kernels.  Eliminating a keyframe's block column is exactly the Schur complement of
// that keyframe onto its <= b successors.  Inputs use the global variable order of include/sage_ba.h
// ([poses | (code, scale)...]); the interleaving is an index map, nothing is copied on the host.
#include "sage_common.cuh"
#include "sage_kernels.h"

namespace sage
{

constexpr int BAND_THREADS = 512;
constexpr int BAND_MAXS = 40; // 7 + C <= 39

__device__ __forceinline__ int band_var(int k, int a, int K, int C)
{
  return a < 6 ? 6 * k + a : 6 * K + k * (C + 1) + (a - 6);
}

// Hd: damped dense matrix (n x n, symmetric, fixed variables already turned into identity rows), gd: gradient.
// band: workspace K * (b+1) * S * S doubles.  delta: solution in the global variable order.  info: 0 ok, k+1 if block k is not SPD.
__global__ void __launch_bounds__(BAND_THREADS, 1)
banded_solve_kernel(const double *__restrict__ Hd, const double *__restrict__ gd, double *__restrict__ band, double *__restrict__ delta,
                    int *__restrict__ info, int n, int K, int C, int b)
{
  extern __shared__ double sm[];
  const int S = 7 + C, SP1 = S + 1;
  double *Lkk = sm;                    // [S][S+1]   current diagonal factor
  double *Lc = Lkk + S * SP1;          // [b][S][S+1] sub-diagonal blocks of the current block column
  double *vec = Lc + (size_t)b * S * SP1; // [S] scratch vector
  __shared__ int bad;
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t BS = (size_t)S * S;
  if (tid == 0)
    bad = 0;
  // gather the band: band[(i*(b+1) + o)] = block (i, i-o), o = 0..b
  for (size_t e = tid; e < (size_t)K * (b + 1) * BS; e += nt)
  {
    const int blk = (int)(e / BS), r = (int)((e % BS) / S), c = (int)(e % S);
    const int i = blk / (b + 1), o = blk % (b + 1), j = i - o;
    band[e] = j >= 0 ? Hd[(size_t)band_var(i, r, K, C) * n + band_var(j, c, K, C)] : 0.0;
  }
  __syncthreads();

  for (int k = 0; k < K; ++k)
  {
    // ---- 1. Cholesky of the diagonal block (in shared memory) ----
    const double *D = band + (size_t)(k * (b + 1)) * BS;
    for (int e = tid; e < S * S; e += nt)
      Lkk[(e / S) * SP1 + e % S] = D[e];
    __syncthreads();
    for (int c = 0; c < S; ++c)
    {
      const double piv = Lkk[c * SP1 + c];
      if (!(piv > 0.0))
      {
        if (tid == 0)
          bad = k + 1;
        break; // uniform: every thread reads the same pivot
      }
      const double d = sqrt(piv);
      __syncthreads();
      for (int r = c + tid; r < S; r += nt)
        Lkk[r * SP1 + c] = (r == c) ? d : Lkk[r * SP1 + c] / d;
      __syncthreads();
      // trailing update of the lower triangle: L[r][q] -= L[r][c] * L[q][c], c < q <= r
      const int m = S - c - 1;
      for (int e = tid; e < m * m; e += nt)
      {
        const int r = c + 1 + e / m, q = c + 1 + e % m;
        if (q <= r)
          Lkk[r * SP1 + q] -= Lkk[r * SP1 + c] * Lkk[q * SP1 + c];
      }
      __syncthreads();
    }
    if (bad)
      break;
    // ---- 2. sub-diagonal blocks: L_ik = A_ik L_kk^-T  (one thread per row: forward substitution) ----
    const int nb = min(b, K - 1 - k);
    for (int e = tid; e < nb * S; e += nt)
    {
      const int o = 1 + e / S, r = e % S;
      const double *X = band + (size_t)((k + o) * (b + 1) + o) * BS + (size_t)r * S; // row r of block (k+o, k)
      double *Y = Lc + (size_t)(o - 1) * S * SP1 + (size_t)r * SP1;
      for (int c = 0; c < S; ++c)
      {
        double s = X[c];
        for (int q = 0; q < c; ++q)
          s -= Y[q] * Lkk[c * SP1 + q];
        Y[c] = s / Lkk[c * SP1 + c];
      }
    }
    __syncthreads();
    // write the factor back (needed by the substitutions) and update the trailing blocks
    {
      double *Dw = band + (size_t)(k * (b + 1)) * BS;
      for (int e = tid; e < S * S; e += nt)
        Dw[e] = (e % S <= e / S) ? Lkk[(e / S) * SP1 + e % S] : 0.0;
      for (int e = tid; e < nb * S * S; e += nt)
      {
        const int o = 1 + e / (S * S), r = (e % (S * S)) / S, c = e % S;
        band[(size_t)((k + o) * (b + 1) + o) * BS + (size_t)r * S + c] = Lc[(size_t)(o - 1) * S * SP1 + r * SP1 + c];
      }
      // A_ij -= L_ik L_jk^T for k < j <= i <= k + nb, 3x3 register tiles (S = 39 = 13 * 3; generic tail handled by bounds)
      const int npair = nb * (nb + 1) / 2;
      const int TS = (S + 2) / 3;
      for (int e = tid; e < npair * TS * TS; e += nt)
      {
        int pr = e / (TS * TS);
        const int r0 = ((e % (TS * TS)) / TS) * 3, c0 = (e % TS) * 3;
        int oi = 1; // pair index -> (oi >= oj)
        while (pr >= oi)
        {
          pr -= oi;
          ++oi;
        }
        const int oj = pr + 1;
        const double *Li = Lc + (size_t)(oi - 1) * S * SP1;
        const double *Lj = Lc + (size_t)(oj - 1) * S * SP1;
        double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        const int ra = min(r0, S - 1), rb = min(r0 + 1, S - 1), rc = min(r0 + 2, S - 1);
        const int ca = min(c0, S - 1), cb = min(c0 + 1, S - 1), cc = min(c0 + 2, S - 1);
        for (int q = 0; q < S; ++q)
        {
          const double a0 = Li[ra * SP1 + q], a1 = Li[rb * SP1 + q], a2 = Li[rc * SP1 + q];
          const double b0 = Lj[ca * SP1 + q], b1 = Lj[cb * SP1 + q], b2 = Lj[cc * SP1 + q];
          acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[0][2] += a0 * b2;
          acc[1][0] += a1 * b0; acc[1][1] += a1 * b1; acc[1][2] += a1 * b2;
          acc[2][0] += a2 * b0; acc[2][1] += a2 * b1; acc[2][2] += a2 * b2;
        }
        double *dstb = band + (size_t)((k + oi) * (b + 1) + (oi - oj)) * BS;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (r0 + i < S && c0 + j < S)
              dstb[(size_t)(r0 + i) * S + c0 + j] -= acc[i][j];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad)
  {
    if (tid == 0)
      *info = bad;
    for (int e = tid; e < n; e += nt)
      delta[e] = 0.0;
    return;
  }
  if (tid == 0)
    *info = 0;

  // ---- forward substitution  L y = g  (y overwrites delta, block by block) ----
  // Off-diagonal contributions: one thread per (row, previous block) partial dot product, summed through shared memory;
  // the S x S triangular solve runs column by column on one warp with the diagonal factor staged in shared memory.
  double *part = Lc; // [b][S] partial sums (Lc is free now)
  for (int k = 0; k < K; ++k)
  {
    const int nb = min(b, k);
    for (int e = tid; e < S * S; e += nt)
      Lkk[(e / S) * SP1 + e % S] = band[(size_t)(k * (b + 1)) * BS + e];
    for (int e = tid; e < nb * S; e += nt)
    {
      const int o = 1 + e / S, r = e % S;
      const double *Lb = band + (size_t)(k * (b + 1) + o) * BS + (size_t)r * S; // block (k, k-o), row r
      double s = 0.0;
      for (int q = 0; q < S; ++q)
        s += Lb[q] * delta[band_var(k - o, q, K, C)];
      part[(o - 1) * S + r] = s;
    }
    __syncthreads();
    for (int r = tid; r < S; r += nt)
    {
      double s = gd[band_var(k, r, K, C)];
      for (int o = 0; o < nb; ++o)
        s -= part[o * S + r];
      vec[r] = s;
    }
    __syncthreads();
    if (tid < 32)
    {
      for (int c = 0; c < S; ++c)
      {
        const double xc = vec[c] / Lkk[c * SP1 + c];
        __syncwarp();
        for (int r = c + 1 + tid; r < S; r += 32)
          vec[r] -= Lkk[r * SP1 + c] * xc;
        if (tid == 0)
          vec[c] = xc;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = tid; r < S; r += nt)
      delta[band_var(k, r, K, C)] = vec[r];
    __syncthreads();
  }
  // ---- backward substitution  L^T x = y ----
  for (int k = K - 1; k >= 0; --k)
  {
    const int nb = min(b, K - 1 - k);
    for (int e = tid; e < S * S; e += nt)
      Lkk[(e / S) * SP1 + e % S] = band[(size_t)(k * (b + 1)) * BS + e];
    for (int e = tid; e < nb * S; e += nt)
    {
      const int o = 1 + e / S, r = e % S;
      const double *Lb = band + (size_t)((k + o) * (b + 1) + o) * BS; // block (k+o, k): use its transpose
      double s = 0.0;
      for (int q = 0; q < S; ++q)
        s += Lb[(size_t)q * S + r] * delta[band_var(k + o, q, K, C)];
      part[(o - 1) * S + r] = s;
    }
    __syncthreads();
    for (int r = tid; r < S; r += nt)
    {
      double s = delta[band_var(k, r, K, C)];
      for (int o = 0; o < nb; ++o)
        s -= part[o * S + r];
      vec[r] = s;
    }
    __syncthreads();
    if (tid < 32)
    {
      for (int c = S - 1; c >= 0; --c)
      {
        const double xc = vec[c] / Lkk[c * SP1 + c];
        __syncwarp();
        for (int r = tid; r < c; r += 32)
          vec[r] -= Lkk[c * SP1 + r] * xc; // (L^T)[r][c] = L[c][r]
        if (tid == 0)
          vec[c] = xc;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = tid; r < S; r += nt)
      delta[band_var(k, r, K, C)] = vec[r];
    __syncthreads();
  }
}

size_t banded_smem_bytes(int C, int b)
{
  const int S = 7 + C;
  return sizeof(double) * ((size_t)S * (S + 1) * (1 + b) + S);
}

size_t banded_workspace_doubles(int K, int C, int b) { return (size_t)K * (b + 1) * (7 + C) * (7 + C); }

int launch_banded_solve(const double *Hd, const double *gd, double *band, double *delta, int *info, int n, int K, int C, int b,
                        cudaStream_t stream)
{
  const size_t smem = banded_smem_bytes(C, b);
  static size_t configured = 0;
  if (smem > configured)
  {
    if (cudaFuncSetAttribute(banded_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return -1;
    configured = smem;
  }
  banded_solve_kernel<<<1, BAND_THREADS, smem, stream>>>(Hd, gd, band, delta, info, n, K, C, b);
  return 0;
}

} // namespace sage

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
      // A_ij -= L_ik L_jk^T for k < j <= i <= k + nb
      const int npair = nb * (nb + 1) / 2;
      for (int e = tid; e < npair * S * S; e += nt)
      {
        int pr = e / (S * S);
        const int r = (e % (S * S)) / S, c = e % S;
        int oi = 1, oj = 1; // pair index -> (oi >= oj)
        while (pr >= oi)
        {
          pr -= oi;
          ++oi;
        }
        oj = pr + 1;
        const double *Li = Lc + (size_t)(oi - 1) * S * SP1 + (size_t)r * SP1;
        const double *Lj = Lc + (size_t)(oj - 1) * S * SP1 + (size_t)c * SP1;
        double s = 0.0;
        for (int q = 0; q < S; ++q)
          s += Li[q] * Lj[q];
        band[(size_t)((k + oi) * (b + 1) + (oi - oj)) * BS + (size_t)r * S + c] -= s;
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
  for (int k = 0; k < K; ++k)
  {
    for (int r = tid; r < S; r += nt)
    {
      double s = gd[band_var(k, r, K, C)];
      for (int o = 1; o <= min(b, k); ++o)
      {
        const double *Lb = band + (size_t)(k * (b + 1) + o) * BS + (size_t)r * S; // block (k, k-o), row r
        for (int q = 0; q < S; ++q)
          s -= Lb[q] * delta[band_var(k - o, q, K, C)];
      }
      vec[r] = s;
    }
    __syncthreads();
    if (tid == 0)
    {
      const double *Lb = band + (size_t)(k * (b + 1)) * BS;
      for (int r = 0; r < S; ++r)
      {
        double s = vec[r];
        for (int q = 0; q < r; ++q)
          s -= Lb[(size_t)r * S + q] * vec[q];
        vec[r] = s / Lb[(size_t)r * S + r];
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
    for (int r = tid; r < S; r += nt)
    {
      double s = delta[band_var(k, r, K, C)];
      for (int o = 1; o <= min(b, K - 1 - k); ++o)
      {
        const double *Lb = band + (size_t)((k + o) * (b + 1) + o) * BS; // block (k+o, k): use its transpose
        for (int q = 0; q < S; ++q)
          s -= Lb[(size_t)q * S + r] * delta[band_var(k + o, q, K, C)];
      }
      vec[r] = s;
    }
    __syncthreads();
    if (tid == 0)
    {
      const double *Lb = band + (size_t)(k * (b + 1)) * BS;
      for (int r = S - 1; r >= 0; --r)
      {
        double s = vec[r];
        for (int q = r + 1; q < S; ++q)
          s -= Lb[(size_t)q * S + r] * vec[q];
        vec[r] = s / Lb[(size_t)r * S + r];
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

// tc_probe.cu -- known-answer tests of the tcgen05 conventions geometric.cu relies on (sm_100a): tensor-memory addressing
// (tcgen05.st / tcgen05.ld round trip), the no-swizzle shared-memory descriptors (K-major and MN-major: which field is the stride
// between 16-byte chunks along MN, which the stride between 8-row K groups), the kind::tf32 instruction descriptor, the
// accumulator layout (lane = row, column = column) and the commit -> mbarrier protocol.  Standalone binary
// (sage-slam_b200/lib/tc_probe, built by build.py; tests/test_gpu_tcgen05.py runs it):
//   tc_probe <test> <variant> [reps]
//   test 0: st -> ld round trip     test 1: MN-major operands     test 2: K-major operands     test 3: K-major, padded strides
//   variant bit 0: swap the two stride fields;  bit 1: pre-fill the accumulator with 7 and accumulate into it
// Small integers are exact in tf32, so the expected J^T J is exact and the comparison is ==.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "tcgen05.cuh"

using namespace sage;

constexpr int NCH = 40; // 16-byte chunks along MN per K group: 160 columns
constexpr int GROUP_BYTES = NCH * 128 + 1024; // room for the padded variant (test 3)
constexpr int G = 16; // K groups of 8 rows
constexpr int PAD_SBO = 272, PAD_LBO = 144, PAD_KSTEP = 20 * PAD_SBO; // geometric.cu's bank-conflict-free K-major staging
constexpr int M = 128, N = 160;

__host__ __device__ inline float probe_value(int k, int m) { return (float)(((k * 37 + m * 11 + (k ^ m)) % 7) - 3); }

__global__ void __launch_bounds__(128) probe_kernel(float *D, float *S, int test, int variant, int reps, long long *cycles)
{
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0)
    tc::tmem_alloc(&tmem_slot, 256);
  if (threadIdx.x == 0)
  {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  for (int e = threadIdx.x; e < G * 8 * N; e += blockDim.x)
  {
    const int k = e / N, m = e % N;
    size_t off;
    if (test == 3) // K-major with padded strides
      off = (size_t)(k / 8) * PAD_KSTEP + (m / 8) * PAD_SBO + ((k % 8) / 4) * PAD_LBO + (m % 8) * 16 + (k % 4) * 4;
    else if (test == 2) // K-major: core matrix = 8 MN rows x 16 bytes (4 K elements); two core matrices along K are adjacent
      off = (size_t)(k / 8) * GROUP_BYTES + (m / 8) * 256 + ((k % 8) / 4) * 128 + (m % 8) * 16 + (k % 4) * 4;
    else // MN-major: core matrix = 8 K rows x 16 bytes (4 MN elements)
      off = (size_t)(k / 8) * GROUP_BYTES + (m / 4) * 128 + (k % 8) * 16 + (m % 4) * 4;
    *reinterpret_cast<float *>(stage + off) = probe_value(k, m);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0)
    cycles[1] = tmem;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (test == 0 || (variant & 2))
  {
    for (int c = 0; c < N; c += 16)
    {
      float v[16];
      for (int i = 0; i < 16; ++i)
        v[i] = test == 0 ? (float)((warp * 32 + lane) * 1000 + c + i) : 7.f;
      tc::tmem_st16(tmem + lane_base + c, v);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  if (test != 0)
  {
    uint32_t f0, f1; // byte strides: f0 -> "stride byte offset" field, f1 -> "leading byte offset" field
    if (test == 1)
      f0 = 128, f1 = GROUP_BYTES; // chunk stride along MN, K-group stride
    else if (test == 3)
      f0 = PAD_SBO, f1 = PAD_LBO;
    else
      f0 = 256, f1 = 128; // 8-row group stride along MN, core-matrix stride along K
    if (variant & 1)
    {
      const uint32_t t = f0;
      f0 = f1;
      f1 = t;
    }
    const uint32_t idesc = tc::idesc_tf32(M, N, test == 1);
    long long t0 = 0;
    uint32_t phase = 0;
    if (threadIdx.x == 0)
      t0 = clock64();
    for (int r = 0; r < reps; ++r)
    {
      if (threadIdx.x == 0)
      {
        for (int g = 0; g < G; ++g)
        {
          const uint64_t d = tc::smem_desc(tc::smem_u32(stage + g * (test == 3 ? PAD_KSTEP : GROUP_BYTES)), f0, f1);
          tc::mma_tf32_ss(tmem, d, d, idesc, (g > 0 || (variant & 2)) ? 1u : 0u);
        }
        tc::mma_commit(&bar);
      }
      tc::mbar_wait(&bar, phase);
      phase ^= 1;
    }
    tc::fence_after_sync();
    if (threadIdx.x == 0)
      cycles[0] = clock64() - t0;
  }
  for (int c = 0; c < N; c += 16)
  {
    float v[16];
    tc::tmem_ld16(tmem + lane_base + c, v);
    for (int i = 0; i < 16; ++i)
      D[(warp * 32 + lane) * N + c + i] = v[i];
  }
  for (int e = threadIdx.x; e < G * GROUP_BYTES / 4; e += blockDim.x) // the stage as the generic proxy sees it
    S[e] = reinterpret_cast<const float *>(stage)[e];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0)
    tc::tmem_dealloc(tmem, 256);
}

int main(int argc, char **argv)
{
  const int test = argc > 1 ? atoi(argv[1]) : 1;
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  const int reps = argc > 3 ? atoi(argv[3]) : 1;
  float *D, *S;
  long long *cyc;
  const int smem = G * GROUP_BYTES;
  cudaMalloc(&D, sizeof(float) * M * N);
  cudaMalloc(&S, smem);
  cudaMalloc(&cyc, 2 * sizeof(long long));
  cudaMemset(D, 0xff, sizeof(float) * M * N);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(D, S, test, variant, reps, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess)
  {
    printf("test %d variant %d: CUDA error %s\n", test, variant, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> h(M * N), hs(smem / 4);
  long long hc[2] = {0, 0};
  cudaMemcpy(h.data(), D, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs.data(), S, smem, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  int bad = 0, bad80 = 0, nz = 0;
  for (float v : hs)
    nz += v != 0.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n)
    {
      float ex = 0.f;
      if (test == 0)
        ex = (float)(m * 1000 + n);
      else
      {
        for (int k = 0; k < G * 8; ++k)
          ex += probe_value(k, m) * probe_value(k, n);
        ex = ex * (float)((variant & 2) ? reps : 1) + ((variant & 2) ? 7.f : 0.f);
      }
      if (h[m * N + n] != ex)
      {
        if (bad < 4)
          printf("  D[%d][%d] = %g expected %g\n", m, n, h[m * N + n], ex);
        ++bad;
        if (m < 80 && n < 80)
          ++bad80;
      }
    }
  printf("test %d variant %d reps %d: tmem base 0x%llx, stage nonzeros %d, mismatches %d of %d (%d in the 80x80 corner)", test, variant, reps,
         hc[1], nz, bad, M * N, bad80);
  if (test != 0)
    printf(", %.1f cycles per M128 N160 K8 mma (issue -> commit -> wait, %d per rep)", (double)hc[0] / (double)(reps * G), G);
  printf("\n  D[0][0..7] =");
  for (int i = 0; i < 8; ++i)
    printf(" %g", h[i]);
  printf("   D[1][0..3] =");
  for (int i = 0; i < 4; ++i)
    printf(" %g", h[N + i]);
  printf("\n");
  return bad ? 1 : 0;
}

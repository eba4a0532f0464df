// tcgen05.cuh -- thin inline-PTX wrappers for the 5th-generation tensor cores of sm_100a: tensor-memory allocation, shared-memory
// matrix descriptors (no swizzle), tcgen05.mma kind::tf32 issued by one thread, commit -> mbarrier, tcgen05.ld.
// Internal; used by geometric.cu (the 160-wide rank-k update of the geometric lineariser) and tc_probe.cu.
#pragma once
#include <cstdint>

namespace sage
{
namespace tc
{

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- tensor memory -------------------------------------------------------------------------------------------------------
// one warp allocates `cols` (power of two >= 32) columns x 128 lanes x 32 bit; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// spin with a bound: a descriptor / protocol mistake must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin)
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(a), "r"(parity)
                 : "memory");
    if (spin > (1u << 24))
      __trap();
  }
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor without swizzle.  K-major operand of 32-bit elements (the only major-ness kind::tf32 executes;
// pinned by tc_probe.cu): a "core matrix" is 8 MN rows x 16 bytes (4 consecutive K elements) = 128 contiguous bytes; the next 8 MN
// rows are `sbo_bytes` further ("stride byte offset"), the next 4 K elements `lbo_bytes` further ("leading byte offset").  Both
// are multiples of 16 and need not be 128 / 256: padded strides are legal.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // [0,14)  start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16; // [16,30) leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32; // [32,46) stride byte offset
  d |= (uint64_t)1 << 46;                            // [46,48) descriptor version (sm_100)
  return d;                                          // base offset 0, lbo mode 0, layout type 0 = no swizzle
}
// Instruction descriptor of kind::tf32, fp32 accumulate, dense, no negation; mn_major: both operands MN-major (else K-major)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool mn_major)
{
  return (1u << 4)                     // [4,6)   D format: f32
         | (2u << 7)                   // [7,10)  A format: tf32
         | (2u << 10)                  // [10,13) B format: tf32
         | ((mn_major ? 1u : 0u) << 15) // [15]    A major
         | ((mn_major ? 1u : 0u) << 16) // [16]    B major
         | ((uint32_t)(N >> 3) << 17)  // [17,23) N / 8
         | ((uint32_t)(M >> 4) << 24); // [24,29) M / 16
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}

// ---- tensor memory -> registers --------------------------------------------------------------------------------------------
// 32 lanes x 16 consecutive columns: thread t of the warp receives lane (base lane + t), columns [col, col + 16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v)
{
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = __uint_as_float(r[i]);
}

// registers -> tensor memory (tests)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
               "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
               "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

} // namespace tc
} // namespace sage

// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the tensor-core paths:
// the harmonic mixer inside the fused audio kernel and the encoder GEMMs.
//
// Operands are written to shared memory by ordinary threads (they are produced on chip: sines of the
// oscillator bank, LayerNorm outputs), so no TMA tensor maps are involved; the canonical no-swizzle
// K-major UMMA layout is used:
//     core matrix = 8 rows x 16 bytes (4 tf32), stored as 128 contiguous bytes
//     element (row r, col k) of a [R x K] operand lives at byte
//         (k/4) * (R/8)*128  +  (r/8) * 128  +  (r%8) * 16  +  (k%4) * 4
//     -> SBO (8-row group stride) = 128 B, LBO (16-byte K-chunk stride) = (R/8)*128 B
// fp32 accuracy comes from the 3xTF32 split: x = hi + lo with hi = x & 0xffffe000 (exactly tf32),
// lo = x - hi (exact in fp32), D += A_hi B_hi + A_lo B_hi + A_hi B_lo.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t nws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- 3xTF32 split
__device__ __forceinline__ float nws_tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float nws_tf32_lo(float x, float hi) { return __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u); }

// byte offset of element (r, k) in the canonical K-major no-swizzle layout of an [R x K] operand
__device__ __host__ __forceinline__ uint32_t nws_umma_offset(int r, int k, int R) {
  return (uint32_t)((k >> 2) * (R >> 3) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

// ---------------------------------------------------------------- descriptors
// shared-memory matrix descriptor, SM100 format (version 1), no swizzle
__device__ __forceinline__ uint64_t nws_umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  return d;         // base offset 0, lbo mode 0, layout type 0 (SWIZZLE_NONE)
}

// instruction descriptor for kind::tf32, fp32 accumulate, K-major A and B, dense
__device__ __host__ __forceinline__ uint32_t nws_umma_idesc_tf32(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (2u << 7) | (2u << 10)       // a_format = b_format = TF32
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void nws_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nws_smem_u32(bar)), "r"(count) : "memory");
}
// one arrival (release at CTA scope: the thread's earlier shared-memory stores are visible to whoever observes the phase)
__device__ __forceinline__ void nws_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nws_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t nws_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(nws_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// bounded wait: returns false on timeout (a wrong descriptor must not hang the GPU box)
__device__ __forceinline__ bool nws_mbar_wait(uint64_t* bar, uint32_t parity, uint32_t max_polls = 1u << 26) {
  for (uint32_t i = 0; i < max_polls; ++i)
    if (nws_mbar_try_wait(bar, parity)) return true;
  return false;
}

// one lane of the (converged) warp, chosen by the hardware: the issuing thread of the MMA warps
__device__ __forceinline__ bool nws_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void nws_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void nws_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void nws_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
// warp-collective; writes the base address of `ncols` (power of two >= 32) columns to *dst_smem
__device__ __forceinline__ void nws_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(nws_smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void nws_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues
__device__ __forceinline__ void nws_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void nws_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(nws_smem_u32(bar))
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread (lane l of warp w) receives row 32*(w%4)+l
__device__ __forceinline__ void nws_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// same load without the wait: issue several, then nws_tmem_wait_ld() once
__device__ __forceinline__ void nws_tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void nws_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Pins 16 registers behind the preceding volatile asm (the wait): the compiler cannot move their uses above it.
__device__ __forceinline__ void nws_reg_fence16(float* v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                    "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
}

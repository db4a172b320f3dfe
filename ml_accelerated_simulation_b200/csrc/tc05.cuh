// Thin inline-PTX wrappers for the Blackwell (sm_100a) features used by the convolution kernels:
// mbarrier, 1-D bulk async copy (UBLKCP), tcgen05 MMA / TMEM alloc / TMEM load, proxy fences.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlsim {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------- mbarrier -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Fast path first: measured on B200 (scripts/umma_mix_bench.cu), mbarrier.test_wait on a completed phase returns in 13
// clocks, mbarrier.try_wait in 187 (it is a potentially-suspending instruction).  In steady state the barriers of the conv
// pipelines are complete when their waiter arrives, and the tensor queue is shallow: every clock the MMA-issuing thread
// spends in a wait is a clock of idle tensor pipe.  Only a barrier that is NOT ready falls through to try_wait.
// Bounded spin: a protocol bug becomes a trap (reported as a launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_test_wait(bar, parity)) return;
#ifndef MLSIM_NO_WATCHDOG
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    if (mbar_try_wait(bar, parity)) return;
  }
  printf("mlsim: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
         smem_u32(bar), parity);
  __trap();
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// ------------------------------- proxies / fences -------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------- bulk async copy (global -> shared), 1-D -------------------------------
// bytes % 16 == 0, both addresses 16-byte aligned; completion is counted on `bar` (complete_tx).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------- TMEM -------------------------------
// Executed by one full warp.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (base_lane + t), columns c..c+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// Store 32 lanes x 32 columns of one 32-bit value (used to zero accumulator slots).
__device__ __forceinline__ void tmem_st32_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st16_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------- UMMA -------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"):
//   element (row, k) of a [rows][16] bf16 slice lives at
//     start + (row/8)*SBO + (row%8)*16 + (k/8)*LBO + (k%8)*2       bytes
// (cute::UMMA::SmemDescriptor, canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version 1 (Blackwell)
  // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B (both K-major), fp32 accumulate, M=128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4)                        // c_format = F32
         | (1u << 7)                      // a_format = BF16
         | (1u << 10)                     // b_format = BF16
         | ((uint32_t)(n >> 3) << 17)     // n_dim
         | ((uint32_t)(128 >> 4) << 24);  // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The 12 accumulate-MMAs one staged input row feeds into one accumulator segment (3 dy shifts x 4 K16 slices),
// issued back to back from ONE asm block so the issuing thread spends ~3 instructions per MMA.
// Descriptor start-address fields advance by compile-time constants (16-byte units):
//   A: +A_K16 per K16 slice, +1 per dy (one pixel = 16 B);   B: +B_STEP per (dy, K16) step.
template <int A_K16, int B_STEP>
__device__ __forceinline__ void umma_bf16_row12(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
#define MLSIM_MMA "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
#define MLSIM_NEXT_K "add.u64 da, da, %4;\n\tadd.u64 db, db, %5;\n\t"
#define MLSIM_NEXT_DY "sub.u64 da, da, %6;\n\tadd.u64 db, db, %5;\n\t"
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "mov.b64 da, %1;\n\t"
      "mov.b64 db, %2;\n\t"
      MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_DY
      MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_DY
      MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA MLSIM_NEXT_K MLSIM_MMA
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "n"(A_K16), "n"(B_STEP), "n"(3 * A_K16 - 1)
      : "memory");
#undef MLSIM_MMA
#undef MLSIM_NEXT_K
#undef MLSIM_NEXT_DY
}

// MMAs [I0, I1) of the same 12 (index i: dy = i / 4, K16 slice = i % 4): lets the issuing thread put its bookkeeping
// (commits of the previous row) between two groups of MMAs, where the tensor queue hides it.
template <int A_K16, int B_STEP, int I0, int I1>
__device__ __forceinline__ void umma_bf16_rowpart(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
#pragma unroll
  for (int i = I0; i < I1; ++i) {
    const uint64_t da = desc_a + (uint64_t)((i & 3) * A_K16 + (i >> 2));
    const uint64_t db = desc_b + (uint64_t)(i * B_STEP);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc)
        : "memory");
  }
}

// Make `bar` track completion of all tcgen05 async ops issued so far by this thread.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

}  // namespace tc
}  // namespace mlsim

// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for operand layouts / start alignments.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_bench scripts/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ml_accelerated_simulation_b200/csrc/tc05.cuh"
using namespace mlsim::tc;

struct Variant { int n; int lbo_a; int a_off; int layout; int lbo_b; int nmma; int two_tmem; int tmem_traffic; int bulk_traffic; int commit_every; int ncommit; };

__global__ void __launch_bounds__(192, 1) k_bench(Variant v, unsigned long long* out, const uint8_t* gsrc) {
  __shared__ volatile int stop_flag;
  __shared__ __align__(8) uint64_t cbar;
  __shared__ __align__(8) uint64_t dummy[4];
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 192) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&cbar, 1); for (int c = 0; c < 4; ++c) mbar_init(&dummy[c], 1); stop_flag = 0; fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t abase = smem_u32(smem) + v.a_off;
    const uint32_t bbase = smem_u32(smem) + 64 * 1024;
    uint64_t hi;
    if (v.layout == 0) hi = ((uint64_t)((128u >> 4) | (1u << 14))) << 32;               // no swizzle, SBO 128
    else hi = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;                  // SW128, SBO 1024
    const uint32_t idesc = umma_idesc_bf16(v.n);
    long long t0 = clock64();
    if (v.layout == 2) {
      // conv-kernel style: 12-MMA asm blocks (3 instructions per MMA on the issuing thread), optional commits
      const uint64_t hi0 = ((uint64_t)((128u >> 4) | (1u << 14))) << 32;
      const uint32_t idesc2 = umma_idesc_bf16(v.n);
      for (int i = 0; i < v.nmma / 12; ++i) {
        const uint32_t st = (uint32_t)(i % 6) * (16640u >> 4);
        const uint64_t da = hi0 | (uint64_t)(((abase) >> 4) + st) | ((uint64_t)(2080u >> 4) << 16);
        const uint64_t db = hi0 | (uint64_t)((bbase) >> 4) | ((uint64_t)((uint32_t)v.lbo_b >> 4) << 16);
        umma_bf16_row12<260, 384>(tmem + (uint32_t)(i % 6) * 64u, da, db, idesc2);
        for (int c = 0; c < v.ncommit; ++c) umma_commit(&dummy[c]);
      }
    } else
    for (int i = 0; i < v.nmma; ++i) {
      const int k16 = i & 3, dy = (i >> 2) % 3;
      uint32_t a_lo, b_lo;
      if (v.layout == 0) {
        a_lo = ((abase + dy * 16 + k16 * 2 * v.lbo_a) >> 4) | ((uint32_t)(v.lbo_a >> 4) << 16);
        b_lo = ((bbase + dy * 8 * v.lbo_b + k16 * 2 * v.lbo_b) >> 4) | ((uint32_t)(v.lbo_b >> 4) << 16);
      } else {
        a_lo = ((abase + dy * 16384 + k16 * 32) >> 4) | (1u << 16);
        b_lo = ((bbase + dy * 24576 + k16 * 32) >> 4) | (1u << 16);
      }
      const uint32_t d = tmem + (v.two_tmem ? ((i >> 2) & 1) * 256 : 0);
      umma_bf16(d, hi | a_lo, hi | b_lo, idesc, 1u);
      if (v.commit_every && (i % v.commit_every) == v.commit_every - 1) {
        for (int c = 0; c < v.ncommit; ++c) umma_commit(&dummy[c]);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
    stop_flag = 1;
  } else if (warp >= 2 && v.tmem_traffic) {
    // epilogue-like traffic on TMEM columns 448..511 (never touched by the MMAs above)
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 448;
    uint32_t r[32]; uint32_t acc = 0;
    while (!stop_flag) {
      tmem_ld32(ta, r); tmem_ld_wait(); acc += r[3];
      tmem_ld32(ta + 32, r); tmem_ld_wait(); acc += r[5];
      tmem_st32_fill(ta, 0u); tmem_st32_fill(ta + 32, 0u); tmem_st_wait();
      __nanosleep(v.tmem_traffic);
    }
    if (acc == 0x12345) out[0] = acc;
  } else if (warp == 1 && v.bulk_traffic && (threadIdx.x & 31) == 0) {
    uint32_t par = 0;
    while (!stop_flag) {
      mbar_arrive_expect_tx(&cbar, 8 * 2080);
      for (int c = 0; c < 8; ++c) bulk_g2s(smem + 140 * 1024 + c * 2080, gsrc + ((clock() & 1023) * 32768) + c * 4096, 2080, &cbar);
      mbar_wait(&cbar, par); par ^= 1;
      __nanosleep(v.bulk_traffic);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 148 * 8);
  uint8_t* gsrc; cudaMalloc(&gsrc, 64 << 20); cudaMemset(gsrc, 0, 64 << 20);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  Variant vs[] = {
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 0},   // as the conv kernel: LBO 2080, dy shifts of 16 B
    {192, 2080, 0, 0, 3072, 4800, 0, 200, 0}, // + epilogue-like TMEM ld/st from 4 warps
    {192, 2080, 0, 0, 3072, 4800, 0, 20, 0},  // + heavier TMEM traffic
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 500}, // + bulk copies into smem (16.6 KB per ~1.5 us)
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 20},  // + heavy bulk copies
    {192, 2080, 0, 0, 3072, 4800, 0, 200, 500},
    {192, 2080, 0, 1, 0, 4800, 0, 0, 0},      // SW128 K-major descriptors (1024-aligned atoms)
    {48, 2080, 0, 0, 768, 4800, 0, 0, 0},
    {192, 2080, 0, 2, 3072, 4800, 0, 0, 0, 0, 0},    // conv-style 12-MMA asm blocks, no commit
    {192, 2080, 0, 2, 3072, 4800, 0, 0, 0, 0, 1},    // + 1 commit per block
    {192, 2080, 0, 2, 3072, 4800, 0, 0, 0, 0, 2},    // + 2 commits per block
    {64, 2080, 0, 2, 3072, 4800, 0, 0, 0, 0, 0},     // N=64 blocks
    {128, 2080, 0, 2, 3072, 4800, 0, 0, 0, 0, 0},    // N=128 blocks
    {48, 2080, 0, 2, 768, 4800, 0, 0, 0, 0, 0},      // N=48 blocks (last layer)
    {192, 2080, 0, 2, 3072, 4800, 0, 200, 500, 0, 2},// + traffic
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 0, 12, 1},   // one commit per 12 MMAs
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 0, 12, 2},   // two commits per 12 MMAs (as the conv kernel)
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 0, 12, 3},
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 0, 4, 1},    // GEMM-like: commit every 4
    {192, 2080, 0, 0, 3072, 4800, 0, 0, 0, 1, 1},    // commit after every MMA
  };
  for (auto& v : vs) {
    for (int grid : {1, 148}) {
      k_bench<<<grid, 192, 180 * 1024>>>(v, d, gsrc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
      printf("N=%3d layout=%d tmem_traffic=%d bulk_traffic=%d commit_every=%d x%d grid=%3d : %.1f cycles/MMA (ideal %.0f)\n", v.n, v.layout,
             v.tmem_traffic, v.bulk_traffic, v.commit_every, v.ncommit, grid, s / grid / v.nmma, 128.0 * v.n / 256);
    }
  }
  // sustained run: is the in-kernel MMA rate power-limited?  (clock64 counts SM cycles; wall clock gives TFLOP/s)
  {
    Variant v = {192, 2080, 0, 0, 3072, 400000, 0, 0, 0, 0, 0};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      for (int i = 0; i < 10; ++i) k_bench<<<148, 192, 180 * 1024>>>(v, d, gsrc);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      const double flop = 10.0 * 148 * (double)v.nmma * 2.0 * 128 * 192 * 16;
      printf("sustained rep %d: %.1f ms, %.1f cycles/MMA, %.0f TFLOP/s, implied SM clock %.0f MHz\n", rep, ms,
             s / 148 / v.nmma, flop / (ms * 1e-3) / 1e12, (s / 148) * 10 / (ms * 1e-3) / 1e6);
    }
  }
  return 0;
}

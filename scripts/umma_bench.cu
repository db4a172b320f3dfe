// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for operand layouts / start alignments.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_bench scripts/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ml_accelerated_simulation_b200/csrc/tc05.cuh"
using namespace mlsim::tc;

struct Variant { int n; int lbo_a; int a_off; int layout; int lbo_b; int nmma; int two_tmem; };

__global__ void __launch_bounds__(128, 1) k_bench(Variant v, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
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
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  Variant vs[] = {
    {192, 2080, 0, 0, 3072, 4800, 0},   // as the conv kernel: LBO 2080, dy shifts of 16 B
    {192, 2048, 0, 0, 3072, 4800, 0},   // aligned LBO (dy shifts still 16 B)
    {64, 2080, 0, 0, 3072, 4800, 0},
    {128, 2080, 0, 0, 3072, 4800, 0},
    {256, 2080, 0, 0, 4096, 4800, 0},
    {192, 2080, 0, 1, 0, 4800, 0},      // SW128 K-major descriptors (1024-aligned atoms)
    {64, 2080, 0, 1, 0, 4800, 0},
    {256, 2080, 0, 1, 0, 4800, 0},
    {192, 2080, 0, 0, 3072, 4800, 1},   // alternate accumulators every 4 MMAs
    {48, 2080, 0, 0, 768, 4800, 0},
    {16, 2080, 0, 0, 768, 4800, 0},
  };
  for (auto& v : vs) {
    for (int grid : {1, 148}) {
      k_bench<<<grid, 128, 180 * 1024>>>(v, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
      printf("N=%3d layout=%d lboA=%d two_tmem=%d grid=%3d : %.1f cycles/MMA (ideal %.0f)\n", v.n, v.layout, v.lbo_a,
             v.two_tmem, grid, s / grid / v.nmma, 128.0 * v.n / 256);
    }
  }
  return 0;
}

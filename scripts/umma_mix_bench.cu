// Micro-benchmark: does interleaving two small MMAs of a different shape (first layer: M128 x N64 x K16, other operands,
// other accumulator) into the hidden layer's stream of 12 x (M128 x N192 x K16) cost more than their own tensor time?
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_mix_bench scripts/umma_mix_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ml_accelerated_simulation_b200/csrc/tc05.cuh"
using namespace mlsim::tc;

// mode 0: 12 x N192 only; 1: + 2 x N64 (A tile LBO 2048, own B, accumulator columns 384..447);
// 2: + 2 x N192 (same shape as the stream, accumulator 320..511); 3: + 2 x N64 issued from ANOTHER thread (warp 1)
// mode 4..7: mode 1 plus four more warps that write a 16.6 KB stage + a 6 KB A tile per `period` clocks with generic
// stores (4: no fence; 5: + fence.proxy.async per thread and row; 6: + tcgen05.ld of 64 columns; 7: all, period 0)
__global__ void __launch_bounds__(256, 1) k_mix(int mode, int rows, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  __shared__ volatile int go;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); go = 0; fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  const uint64_t hi0 = ((uint64_t)((128u >> 4) | (1u << 14))) << 32;
  const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem) + 100 * 1024, a0 = smem_u32(smem) + 176 * 1024, b0 = smem_u32(smem) + 184 * 1024;
  auto small = [&](int n, uint32_t col) {
    for (int k16 = 0; k16 < 2; ++k16) {
      const uint64_t da = umma_desc(a0 + k16 * 2 * 2048, 2048, 128);
      const uint64_t db = umma_desc(b0 + k16 * 2 * 1024, 1024, 128);
      umma_bf16(tmem + col, da, db, umma_idesc_bf16(n), k16 > 0);
    }
  };
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < rows; ++i) {
      const uint32_t st = (uint32_t)(i % 6) * (16640u >> 4);
      const uint64_t da = hi0 | (uint64_t)((abase >> 4) + st) | ((uint64_t)(2080u >> 4) << 16);
      const uint64_t db = hi0 | (uint64_t)(bbase >> 4) | ((uint64_t)(3072u >> 4) << 16);
      umma_bf16_row12<260, 384>(tmem + (uint32_t)(i % 4) * 64u, da, db, umma_idesc_bf16(192));
      if (mode == 1 || mode >= 4) small(64, 384 + (i & 1) * 64);
      if (mode == 2) small(192, 320);
      if (mode == 3) go = i + 1;
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    go = -1;
  } else if (warp >= 4 && mode >= 4) {
    const int t = threadIdx.x - 128;
    uint32_t acc = 0;
    while (go >= 0) {
      const long long t0 = clock64();
      uint8_t* dst = smem + 120 * 1024;                   // a scratch "stage" away from the MMA operands
      if (mode >= 6) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 448, r); tmem_ld_wait(); acc += r[1];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 480, r); tmem_ld_wait(); acc += r[2];
      }
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + c * 2080 + (t + 1) * 16) = make_uint4(acc, 1u, 2u, 3u);
      if (mode >= 5) fence_proxy_async_smem();
      for (int c = 0; c < 3; ++c) *reinterpret_cast<uint4*>(dst + 20 * 1024 + c * 2048 + t * 16) = make_uint4(acc, 1u, 2u, 3u);
      if (mode >= 5) fence_proxy_async_smem();
      __syncwarp();
      if (mode != 7) while (clock64() - t0 < 1300 && go >= 0) { }
    }
    if (acc == 0x12345) out[0] = acc;
  } else if (threadIdx.x == 32 && mode == 3) {
    int done = 0;
    while (true) {
      const int g = go;
      if (g < 0) break;
      if (g > done) { small(64, 384 + (done & 1) * 64); ++done; }
    }
    umma_commit(&bar2);
    mbar_wait(&bar2, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// latencies of the synchronisation primitives on the MMA warp's per-row path (one thread, everything already complete)
__global__ void k_lat(unsigned long long* out) {
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tslot;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tslot, 32); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_arrive(&bar[0]);                                   // phase 0 of bar[0] complete
    long long t0, t1;
    const int N = 64;
    t0 = clock64(); for (int i = 0; i < N; ++i) { asm volatile("" ::: "memory"); (void)clock64(); } t1 = clock64(); out[0] = (t1 - t0) / N;
    t0 = clock64(); for (int i = 0; i < N; ++i) mbar_wait(&bar[0], 0); t1 = clock64(); out[1] = (t1 - t0) / N;       // try_wait, ready
    t0 = clock64(); for (int i = 0; i < N; ++i) tc_fence_after(); t1 = clock64(); out[2] = (t1 - t0) / N;
    t0 = clock64(); for (int i = 0; i < N; ++i) { umma_commit(&bar[1]); } t1 = clock64(); out[3] = (t1 - t0) / N;    // commit with nothing pending
    uint32_t ok = 0;
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
      uint32_t r;
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(r) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
      ok += r;
    }
    t1 = clock64(); out[4] = (t1 - t0) / N; out[5] = ok;
    t0 = clock64(); for (int i = 0; i < N; ++i) fence_proxy_async_smem(); t1 = clock64(); out[6] = (t1 - t0) / N;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tslot, 32);
}

int main() {
  {
    unsigned long long* dl; cudaMalloc(&dl, 64);
    k_lat<<<1, 64>>>(dl);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k_lat error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[8]; cudaMemcpy(h, dl, 56, cudaMemcpyDeviceToHost);
    printf("latency (clk): clock64 %llu | mbarrier.try_wait (ready) %llu | tcgen05.fence::after %llu | tcgen05.commit %llu | mbarrier.test_wait (ready) %llu (%llu ok) | fence.proxy.async %llu\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  }
  unsigned long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rows = 2000;
  const char* names[] = {"12 x N192", "12 x N192 + 2 x N64 (same thread)", "12 x N192 + 2 x N192 (same thread)", "12 x N192 + 2 x N64 (other thread)",
                         "+ 4 warps: 22.6 KB of st.shared per row", "+ fence.proxy.async", "+ tcgen05.ld", "all, back to back"};
  for (int mode = 0; mode < 8; ++mode)
    for (int grid : {1, 148}) {
      k_mix<<<grid, 256, 192 * 1024>>>(mode, rows, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
      printf("%-40s grid=%3d : %.0f cycles per row\n", names[mode], grid, s / grid / rows);
    }
  return 0;
}

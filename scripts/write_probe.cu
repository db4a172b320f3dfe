// Write-pattern probe for the first conv layer's epilogue: the same CTA / thread -> address mapping (128 threads per
// tile, 8 stores of 16 B per thread at plane stride), no loads, no MMA.  Build: nvcc -O3 -arch=sm_100a write_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128) k_pattern(uint4* out, int nx, int ny, int batch, int mode) {
  const int nyt = ny / 128, ntiles = batch * nx * nyt;
  const int per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(t0 + per, ntiles);
  const size_t plane = ny + 2;
  for (int t = t0; t < t1; ++t) {
    const int yt = t % nyt, x = (t / nyt) % nx, b = t / (nyt * nx);
    const int y = yt * 128 + threadIdx.x;
    if (mode == 0) {                                    // activation layout: 8 planes, 16 B per pixel per plane
      uint4* row = out + (((size_t)b * nx + x) * 8) * plane + (y + 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) row[(size_t)c * plane] = make_uint4(t, c, y, 1);
    } else if (mode == 2) {                             // planes padded so that every 2 KB chunk starts on a 128-byte line
      const size_t plane2 = ny + 16;
      uint4* row = out + (((size_t)b * nx + x) * 8) * plane2 + (y + 8);
#pragma unroll
      for (int c = 0; c < 8; ++c) row[(size_t)c * plane2] = make_uint4(t, c, y, 1);
    } else {                                            // same bytes, one contiguous 16 KB block per tile
      uint4* blk = out + (size_t)t * 1024;
#pragma unroll
      for (int c = 0; c < 8; ++c) blk[c * 128 + threadIdx.x] = make_uint4(t, c, y, 1);
    }
  }
}
int main() {
  const int nx = 256, ny = 256, batch = 64;
  const size_t n16 = (size_t)batch * nx * 8 * (ny + 16);
  uint4* buf; cudaMalloc(&buf, n16 * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode)
    for (int grid : {148, 592, 4736}) {
      for (int i = 0; i < 3; ++i) k_pattern<<<grid, 128>>>(buf, nx, ny, batch, mode);
      cudaEventRecord(e0);
      for (int i = 0; i < 10; ++i) k_pattern<<<grid, 128>>>(buf, nx, ny, batch, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
      printf("mode %d grid %5d: %.4f ms -> %.2f TB/s\n", mode, grid, ms, (double)batch * nx * ny * 128 / ms / 1e9);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

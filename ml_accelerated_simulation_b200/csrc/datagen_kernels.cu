// Data-generation side of the rollout path: the staggered-velocity downsampler that sits
// between the fine rollout and the paired coarse step in reference src/data_generation.py:91-111,208-212.
//
//   downsample_staggered_velocity_component(u, direction, factor):
//     w   = u sliced along `direction` at indices factor-1, 2*factor-1, ...      (the face that coincides with a coarse face)
//     out = block mean of w over blocks of `factor` along the OTHER axis
//   u (offset (1, 1/2)) uses direction 0 (x = array axis 0), v (offset (1/2, 1)) direction 1 (y, contiguous).
//
// HBM-bound byte shuffling: u reads 1/factor of its rows (full, coalesced float4), v reads one float per `factor`
// columns of every row (one 32-byte sector per value: inherent to the layout); outputs are 1/factor^2 of the input.
#include "common.cuh"

namespace mlsim {

// cu[b][I][J] = mean_{j in [J f, (J+1) f)} u[b][I f + f - 1][j]        one thread per output, f contiguous floats
// cv[b][I][J] = mean_{i in [I f, (I+1) f)} v[b][i][J f + f - 1]        one thread per output, f rows strided
__global__ void __launch_bounds__(256)
k_downsample_staggered(const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ cu,
                       float* __restrict__ cv, int batch, int nx, int ny, int f) {
  pdl_launch_dependents();
  pdl_wait();
  const int cnx = nx / f, cny = ny / f;
  const size_t nout = (size_t)batch * cnx * cny;
  const float inv = 1.0f / (float)f;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < 2 * nout; o += (size_t)gridDim.x * blockDim.x) {
    const bool is_v = o >= nout;
    const size_t e = is_v ? o - nout : o;
    const int J = (int)(e % cny);
    const int I = (int)((e / cny) % cnx);
    const size_t b = e / ((size_t)cny * cnx);
    float s = 0.0f;
    if (!is_v) {
      const float* row = u + (b * nx + (size_t)I * f + f - 1) * ny + (size_t)J * f;
      if ((f & 3) == 0) {
        for (int k = 0; k < f; k += 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(row + k));
          s += (q.x + q.y) + (q.z + q.w);
        }
      } else {
        for (int k = 0; k < f; ++k) s += __ldg(row + k);
      }
      cu[e] = s * inv;
    } else {
      const float* col = v + (b * nx + (size_t)I * f) * ny + (size_t)J * f + f - 1;
      for (int k = 0; k < f; ++k) s += __ldg(col + (size_t)k * ny);
      cv[e] = s * inv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Slab decomposition: push this rank's edge rows into the neighbours' halo rows (peer memory over NVLink).
//   top `lo` rows  [nxl-lo, nxl) -> UPPER neighbour's rows [-lo, 0)         bottom `hi` rows [0, hi) -> LOWER neighbour's
//   rows [nxl, nxl+hi).  Field pair layout uv[2][batch][ext_rows][ny], local row i at ext row i + halo; float4 copies.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_halo_push(const float* __restrict__ uv, float* __restrict__ up_uv, float* __restrict__ dn_uv, int images, int ext_rows,
            int ny, int nxl, int halo, int lo, int hi) {
  pdl_launch_dependents();
  pdl_wait();
  const int ny4 = ny >> 2;
  const size_t per_img_up = (size_t)lo * ny4, per_img_dn = (size_t)hi * ny4;
  const size_t n_up = up_uv ? per_img_up * images : 0, n_dn = dn_uv ? per_img_dn * images : 0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n_up + n_dn; e += (size_t)gridDim.x * blockDim.x) {
    if (e < n_up) {
      const size_t img = e / per_img_up, r = e - img * per_img_up;                 // r = row * ny4 + col4
      const size_t row = r / ny4, c4 = r - row * ny4;
      const float4 val = reinterpret_cast<const float4*>(uv + (img * ext_rows + halo + nxl - lo + row) * ny)[c4];
      reinterpret_cast<float4*>(up_uv + (img * ext_rows + halo - lo + row) * ny)[c4] = val;
    } else {
      const size_t f = e - n_up;
      const size_t img = f / per_img_dn, r = f - img * per_img_dn;
      const size_t row = r / ny4, c4 = r - row * ny4;
      const float4 val = reinterpret_cast<const float4*>(uv + (img * ext_rows + halo + row) * ny)[c4];
      reinterpret_cast<float4*>(dn_uv + (img * ext_rows + halo + nxl + row) * ny)[c4] = val;
    }
  }
}

int launch_halo_push(const float* uv, float* up_uv, float* dn_uv, int images, int ext_rows, int ny, int nxl, int halo,
                     int lo, int hi, cudaStream_t st) {
  const size_t n = (size_t)images * (lo + hi) * (ny / 4);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  MLSIM_CUDA_CHECK(launch_pdl(k_halo_push, dim3((unsigned)blocks), dim3(256), 0, st, uv, up_uv, dn_uv, images, ext_rows, ny,
                              nxl, halo, lo, hi));
  return MLSIM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Initial-condition normalisation (torch_cfd filtered_velocity_field): per trajectory
//   s = max sqrt(u^2 + v^2);  u, v *= max_velocity / s.      Two passes: warp-shuffle + atomicMax reduction, rescale.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_speed_max(const float* __restrict__ u, const float* __restrict__ v, unsigned int* __restrict__ smax2, int n_per_img) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const float4* pu = reinterpret_cast<const float4*>(u + (size_t)b * n_per_img);
  const float4* pv = reinterpret_cast<const float4*>(v + (size_t)b * n_per_img);
  float m = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img / 4; i += gridDim.x * blockDim.x) {
    const float4 a = __ldg(pu + i), c = __ldg(pv + i);
    m = fmaxf(m, fmaxf(fmaxf(a.x * a.x + c.x * c.x, a.y * a.y + c.y * c.y), fmaxf(a.z * a.z + c.z * c.z, a.w * a.w + c.w * c.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&smax2[b], __float_as_uint(m));      // non-negative floats order like their bit patterns
}

__global__ void __launch_bounds__(256)
k_speed_rescale(float* __restrict__ u, float* __restrict__ v, const unsigned int* __restrict__ smax2, int n_per_img,
                float max_velocity) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const float s2 = __uint_as_float(smax2[b]);
  const float f = s2 > 0.0f ? max_velocity * rsqrtf(s2) : 1.0f;
  float4* pu = reinterpret_cast<float4*>(u + (size_t)b * n_per_img);
  float4* pv = reinterpret_cast<float4*>(v + (size_t)b * n_per_img);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img / 4; i += gridDim.x * blockDim.x) {
    float4 a = pu[i], c = pv[i];
    a.x *= f; a.y *= f; a.z *= f; a.w *= f;
    c.x *= f; c.y *= f; c.z *= f; c.w *= f;
    pu[i] = a; pv[i] = c;
  }
}

int launch_normalize_speed(float* u, float* v, unsigned int* scratch, int batch, int n_per_img, float max_velocity,
                           cudaStream_t st) {
  MLSIM_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)batch * sizeof(unsigned int), st));
  int bx = (n_per_img / 4 + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)batch);
  MLSIM_CUDA_CHECK(launch_pdl(k_speed_max, grid, dim3(256), 0, st, (const float*)u, (const float*)v, scratch, n_per_img));
  MLSIM_CUDA_CHECK(launch_pdl(k_speed_rescale, grid, dim3(256), 0, st, u, v, (const unsigned int*)scratch, n_per_img, max_velocity));
  return MLSIM_OK;
}

int launch_downsample_staggered(const float* u, const float* v, float* cu, float* cv, int batch, int nx, int ny, int f,
                                cudaStream_t st) {
  const size_t nout = 2 * (size_t)batch * (nx / f) * (ny / f);
  size_t blocks = (nout + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  MLSIM_CUDA_CHECK(launch_pdl(k_downsample_staggered, dim3((unsigned)blocks), dim3(256), 0, st, u, v, cu, cv, batch, nx, ny, f));
  return MLSIM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Divergence guard (reference test/sim_test.py:82-87 tests isnan on the host after every step): scan (u, v) for
// non-finite values and record the smallest step index at which one was present.  One float4 pass over the state
// (8 B/cell) every `every` steps; the flag is a device word that the caller copies to pinned host memory
// asynchronously, so the step path never synchronises.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_finite_check(const float4* __restrict__ u, const float4* __restrict__ v, size_t n4, unsigned long long step,
               unsigned long long* __restrict__ flag) {
  pdl_launch_dependents();
  pdl_wait();
  bool bad = false;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(u + e), b = __ldg(v + e);
    // x - x is 0 for finite x and NaN for +-inf / NaN
    const float t = (a.x - a.x) + (a.y - a.y) + (a.z - a.z) + (a.w - a.w) + (b.x - b.x) + (b.y - b.y) + (b.z - b.z) + (b.w - b.w);
    bad |= !(t == 0.0f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicMin(flag, step);
}

int launch_finite_check(const float* u, const float* v, size_t n, unsigned long long step, unsigned long long* flag,
                        cudaStream_t st) {
  const size_t n4 = n / 4;                       // field sizes are multiples of 4 (ny is a power of two >= 64)
  const size_t want = (n4 + 255) / 256;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  MLSIM_CUDA_CHECK(launch_pdl(k_finite_check, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float4*>(u),
                              reinterpret_cast<const float4*>(v), n4, step, flag));
  return MLSIM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// fp64 boundary of the drop-in route: the reference scripts hold their state in float64
// (torch.set_default_dtype(torch.float64), src/inference.py:60); the step arithmetic is fp32.  One fused pass each
// way instead of separate dtype casts, copies and allocations per field.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cast_f64_to_f32(const double2* __restrict__ a, const double2* __restrict__ b, float2* __restrict__ oa,
                  float2* __restrict__ ob, size_t n2) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2; e += (size_t)gridDim.x * blockDim.x) {
    const double2 x = __ldg(a + e), y = __ldg(b + e);
    oa[e] = make_float2((float)x.x, (float)x.y);
    ob[e] = make_float2((float)y.x, (float)y.y);
  }
}

__global__ void __launch_bounds__(256)
k_cast_f32_to_f64(const float2* __restrict__ a, const float2* __restrict__ b, const float2* __restrict__ c,
                  double2* __restrict__ oa, double2* __restrict__ ob, double2* __restrict__ oc, size_t n2) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2; e += (size_t)gridDim.x * blockDim.x) {
    const float2 x = a[e], y = b[e];
    oa[e] = make_double2((double)x.x, (double)x.y);
    ob[e] = make_double2((double)y.x, (double)y.y);
    if (oc != nullptr) {
      const float2 z = c[e];
      oc[e] = make_double2((double)z.x, (double)z.y);
    }
  }
}

static int cast_grid(size_t n2) {
  const size_t want = (n2 + 255) / 256;
  return (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
}
int launch_cast_in(const double* a, const double* b, float* oa, float* ob, size_t n, cudaStream_t st) {
  MLSIM_CUDA_CHECK(launch_pdl(k_cast_f64_to_f32, dim3(cast_grid(n / 2)), dim3(256), 0, st,
                              reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b),
                              reinterpret_cast<float2*>(oa), reinterpret_cast<float2*>(ob), n / 2));
  return MLSIM_OK;
}
int launch_cast_out(const float* a, const float* b, const float* c, double* oa, double* ob, double* oc, size_t n,
                    cudaStream_t st) {
  MLSIM_CUDA_CHECK(launch_pdl(k_cast_f32_to_f64, dim3(cast_grid(n / 2)), dim3(256), 0, st,
                              reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b),
                              reinterpret_cast<const float2*>(c), reinterpret_cast<double2*>(oa),
                              reinterpret_cast<double2*>(ob), reinterpret_cast<double2*>(oc), n / 2));
  return MLSIM_OK;
}

}  // namespace mlsim

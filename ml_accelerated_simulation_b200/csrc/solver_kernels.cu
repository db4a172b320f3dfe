// Solver half of the rollout step: explicit terms + RK combine (K1), divergence + row FFT (K2),
// column FFT * inverse Laplacian spectrum * column iFFT (K3), row iFFT + gradient subtraction (K4).
//
// Stands in for torch_cfd.fvm.RKStepper.forward / NavierStokes2DFVMProjection as called at
// reference src/inference.py:101 (algorithm: SURVEY.md App. A.2-A.4).  All kernels are HBM-bound:
// state fields are fp32 [batch][nx][ny] (y contiguous); the spectral buffer is float2
// [batch][nx][nycp] with nyc = ny/2+1 valid columns.
#include "common.cuh"
#include "fft.cuh"
#include "solver_kernels.h"

#include <cooperative_groups.h>
#include <stdlib.h>

namespace mlsim {

// ---------------------------------------------------------------------------------------------
// K1: explicit terms F(u,v) and Runge-Kutta stage combine.
// Tile: TX=32 rows (x) by TY columns (y); each thread owns 4 consecutive columns and marches over
// RPG rows so the x-face flux is reused from the previous row; y-face fluxes are shared between
// the thread's four columns (5 evaluations for 4 cells).
// ---------------------------------------------------------------------------------------------
constexpr int K1_TX = 32;
constexpr int K1_HALO = 2;
constexpr int K1_THREADS = 256;

__device__ __forceinline__ float rcp_approx(float x) {        // bare MUFU.RCP (no range fix-up code around it)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int ADV>
__device__ __forceinline__ float face_flux(float cl, float c, float cr, float crr, float w2, float hq) {
  // TWICE the TVD face flux (App. A.3.2):  face = low - (low - high) * phi,  flux = face * w.
  // w2 = sum of the two velocities whose mean is the face velocity (w = w2/2; the caller folds the 1/2 into 1/h);
  // hq = 0.5*dt/h, so the Courant number is C = hq*w2.  With d = cr - c and t2 = +-1 - C:  high = low + (t2/2)*d.
  // van Leer: phi(r) = 2r/(1+r), r = num/d  =>  (t2/2)*phi*d = t2 * (num*d)/(d+num): one product e = num*d whose
  // sign is also the "r > 0" test, one reciprocal (MUFU.RCP, <= 2 ulp; e > 0 implies |d+num| > 0), all selects --
  // no divergent branches.  (If e underflows, both differences are < 1e-19 and the dropped correction < 1e-38.)
  const bool pos = w2 > 0.0f;
  const float low = pos ? c : cr;
  if (ADV == MLSIM_ADV_UPWIND) return low * w2;
  const float d = cr - c;
  const float t2 = fmaf(-hq, w2, pos ? 1.0f : -1.0f);
  if (ADV == MLSIM_ADV_LAX_WENDROFF) return fmaf(t2, 0.5f * d, low) * w2;
  const float num = pos ? (c - cl) : (crr - cr);
  const float e = num * d;
  const float g = e * rcp_approx(d + num);
  return fmaf(t2, (e > 0.0f) ? g : 0.0f, low) * w2;
}

// Same flux from precomputed differences (dl = c - cl, d = cr - c, dr = crr - cr): neighbouring faces share them and the
// Laplacian is d - dl, so the stencil loads every difference once.
template <int ADV>
__device__ __forceinline__ float face_flux_d(float c, float cr, float dl, float d, float dr, float w2, float hq) {
  const bool pos = w2 > 0.0f;
  const float low = pos ? c : cr;
  if (ADV == MLSIM_ADV_UPWIND) return low * w2;
  const float t2 = fmaf(-hq, w2, pos ? 1.0f : -1.0f);
  if (ADV == MLSIM_ADV_LAX_WENDROFF) return fmaf(t2, 0.5f * d, low) * w2;
  const float num = pos ? dl : dr;
  const float e = num * d;
  const float g = e * rcp_approx(d + num);
  return fmaf(t2, (e > 0.0f) ? g : 0.0f, low) * w2;
}

// SLAB (x-slab decomposition): no periodic wrap in x; the tile rows come from the halo-extended buffers (row index
// clamped to the valid rows [-halo, nx+halo)), and one extra leading tile (blockIdx.y == 0 <-> rows [-32, 0)) computes
// output row -1 -- the x-halo of u* that the divergence of the local row 0 needs -- so no second exchange is required.
template <int TY, int ADV, bool SLAB, int MINB>
__global__ void __launch_bounds__(K1_THREADS, MINB)
k_explicit_rk(StageArgs a, SolverParams p, const float* __restrict__ forcing) {
  constexpr int NTY = TY / 4;                 // threads along y
  constexpr int NG = K1_THREADS / NTY;        // row groups
  constexpr int RPG = K1_TX / NG;             // rows marched per thread
  constexpr int SW = TY + 8;                  // smem row width (4-column halo each side, float4 aligned)
  constexpr int SH = K1_TX + 2 * K1_HALO;
  extern __shared__ float4 smem4[];
  float* su = reinterpret_cast<float*>(smem4);
  float* sv = su + SH * SW;

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z;
  const int i0 = SLAB ? ((int)blockIdx.y - 1) * K1_TX : blockIdx.y * K1_TX;
  const int j0 = blockIdx.x * TY;
  const ptrdiff_t img = (ptrdiff_t)b * p.img_stride;
  const float* __restrict__ gu = a.us + img;
  const float* __restrict__ gv = a.vs + img;

  // ---- load the (TX+4) x (TY+8) halo tile, periodic wrap, float4 granules ----
  constexpr int W4 = SW / 4;
  const int ny4 = p.ny >> 2;
  // all of a thread's granules are requested before the first one is stored to shared memory (ncu at 1024^2 x 16: a
  // quarter of K1's stall samples sat on the store that followed each pair of loads)
  {
    constexpr int NIT = (SH * W4 + K1_THREADS - 1) / K1_THREADS;
    float4 tu[NIT], tv[NIT];
#pragma unroll
    for (int q = 0; q < NIT; ++q) {
      const int idx = threadIdx.x + q * K1_THREADS;
      if (idx < SH * W4) {
        const int r = idx / W4, c4 = idx - r * W4;
        const int gi = SLAB ? max(-p.halo, min(i0 - K1_HALO + r, p.nx + p.halo - 1)) : ((i0 - K1_HALO + r) & (p.nx - 1));
        const int gc = ((j0 >> 2) - 1 + c4) & (ny4 - 1);
        tu[q] = __ldg(reinterpret_cast<const float4*>(gu + (ptrdiff_t)gi * p.ny) + gc);
        tv[q] = __ldg(reinterpret_cast<const float4*>(gv + (ptrdiff_t)gi * p.ny) + gc);
      }
    }
#pragma unroll
    for (int q = 0; q < NIT; ++q) {
      const int idx = threadIdx.x + q * K1_THREADS;
      if (idx < SH * W4) {
        const int r = idx / W4, c4 = idx - r * W4;
        reinterpret_cast<float4*>(su + r * SW)[c4] = tu[q];
        reinterpret_cast<float4*>(sv + r * SW)[c4] = tv[q];
      }
    }
  }
  __syncthreads();

  const int ty = threadIdx.x % NTY;
  const int g = threadIdx.x / NTY;
  const int lj = 4 + 4 * ty;                  // smem column of the thread's first cell
  const int li0 = K1_HALO + g * RPG;          // smem row of the thread's first cell

#define SU(i, j) su[(i) * SW + (j)]
#define SV(i, j) sv[(i) * SW + (j)]

  const float hcx = 0.5f * p.dt_hx, hcy = 0.5f * p.dt_hy;     // hq of face_flux
  const float hix = 0.5f * p.inv_hx, hiy = 0.5f * p.inv_hy;    // fluxes carry a factor 2
  // x-face fluxes at face (li0-1) for the 4 columns: carried along the march
  float fxu_prev[4], fxv_prev[4];
  {
    const int i = li0 - 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lj + k;
      const float wxu = SU(i, j) + SU(i + 1, j);
      fxu_prev[k] = face_flux<ADV>(SU(i - 1, j), SU(i, j), SU(i + 1, j), SU(i + 2, j), wxu, hcx);
      const float wxv = SU(i, j) + SU(i, j + 1);
      fxv_prev[k] = face_flux<ADV>(SV(i - 1, j), SV(i, j), SV(i + 1, j), SV(i + 2, j), wxv, hcx);
    }
  }

  float frc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) frc[k] = __ldg(&forcing[j0 + 4 * ty + k]);

#pragma unroll 1
  for (int rr = 0; rr < RPG; ++rr) {
    const int i = li0 + rr;
    // issue this row's RK operand loads first so their latency hides behind the flux arithmetic
    const ptrdiff_t off = img + (ptrdiff_t)(i0 + (i - K1_HALO)) * p.ny + (j0 + 4 * ty);
    const bool live = !SLAB || (i0 + (i - K1_HALO) >= -1);     // slab: the leading tile only produces row -1
    float4 bu = make_float4(0.f, 0.f, 0.f, 0.f), bv = bu, au = bu, av = bu;
    if (live && a.mode == 2) {
      bu = __ldg(reinterpret_cast<const float4*>(a.u0 + off));
      bv = __ldg(reinterpret_cast<const float4*>(a.v0 + off));
    }
    if (live && a.mode >= 2) {
      au = *reinterpret_cast<const float4*>(a.accu + off);
      av = *reinterpret_cast<const float4*>(a.accv + off);
    }
    float fxu[4], fxv[4], fyu[5], fyv[5], lapu[4], lapv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lj + k;
      // x-direction differences of this column: (i-1 -> i), (i -> i+1), (i+1 -> i+2)
      const float um = SU(i - 1, j), u0 = SU(i, j), u1 = SU(i + 1, j), u2 = SU(i + 2, j);
      const float vm = SV(i - 1, j), v0 = SV(i, j), v1 = SV(i + 1, j), v2 = SV(i + 2, j);
      const float dum = u0 - um, du0 = u1 - u0, dup = u2 - u1;
      const float dvm = v0 - vm, dv0 = v1 - v0, dvp = v2 - v1;
      const float wxu = u0 + u1;
      fxu[k] = face_flux_d<ADV>(u0, u1, dum, du0, dup, wxu, hcx);
      const float wxv = u0 + SU(i, j + 1);
      fxv[k] = face_flux_d<ADV>(v0, v1, dvm, dv0, dvp, wxv, hcx);
      lapu[k] = (du0 - dum) * p.nu_hx2;
      lapv[k] = (dv0 - dvm) * p.nu_hx2;
    }
    {
      // y-direction differences of the row: columns lj-2 .. lj+5 -> 7 differences shared by the 5 faces and 4 Laplacians
      float dyu[7], dyv[7];
#pragma unroll
      for (int m = 0; m < 7; ++m) {
        dyu[m] = SU(i, lj - 1 + m) - SU(i, lj - 2 + m);
        dyv[m] = SV(i, lj - 1 + m) - SV(i, lj - 2 + m);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int j = lj - 1 + k;               // y-face between columns j and j+1
        const float wyu = SV(i, j) + SV(i + 1, j);
        fyu[k] = face_flux_d<ADV>(SU(i, j), SU(i, j + 1), dyu[k], dyu[k + 1], dyu[k + 2], wyu, hcy);
        const float wyv = SV(i, j) + SV(i, j + 1);
        fyv[k] = face_flux_d<ADV>(SV(i, j), SV(i, j + 1), dyv[k], dyv[k + 1], dyv[k + 2], wyv, hcy);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lapu[k] = fmaf(dyu[k + 2] - dyu[k + 1], p.nu_hy2, lapu[k]);
        lapv[k] = fmaf(dyv[k + 2] - dyv[k + 1], p.nu_hy2, lapv[k]);
      }
    }
    float ku[4], kv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lj + k;
      const float cu = SU(i, j), cv = SV(i, j);
      const float advu = -((fxu[k] - fxu_prev[k]) * hix + (fyu[k + 1] - fyu[k]) * hiy);
      const float advv = -((fxv[k] - fxv_prev[k]) * hix + (fyv[k + 1] - fyv[k]) * hiy);
      ku[k] = advu + lapu[k] - p.drag * cu + frc[k];
      kv[k] = advv + lapv[k] - p.drag * cv;
      fxu_prev[k] = fxu[k];
      fxv_prev[k] = fxv[k];
    }
    if (!live) continue;
    // ---- RK combine + float4 I/O (the u0 / acc loads were issued at the top of the row) ----
    float4 ou, ov;
    if (a.mode == 0) {
      ou = make_float4(ku[0], ku[1], ku[2], ku[3]);
      ov = make_float4(kv[0], kv[1], kv[2], kv[3]);
    } else if (a.mode == 3) {
      ou = make_float4(fmaf(a.cs, ku[0], au.x), fmaf(a.cs, ku[1], au.y), fmaf(a.cs, ku[2], au.z), fmaf(a.cs, ku[3], au.w));
      ov = make_float4(fmaf(a.cs, kv[0], av.x), fmaf(a.cs, kv[1], av.y), fmaf(a.cs, kv[2], av.z), fmaf(a.cs, kv[3], av.w));
    } else {
      if (a.mode == 1) {        // u_s == u0: already in shared memory
        bu = make_float4(SU(i, lj), SU(i, lj + 1), SU(i, lj + 2), SU(i, lj + 3));
        bv = make_float4(SV(i, lj), SV(i, lj + 1), SV(i, lj + 2), SV(i, lj + 3));
        au = bu;
        av = bv;
      }
      au = make_float4(fmaf(a.ca, ku[0], au.x), fmaf(a.ca, ku[1], au.y), fmaf(a.ca, ku[2], au.z), fmaf(a.ca, ku[3], au.w));
      av = make_float4(fmaf(a.ca, kv[0], av.x), fmaf(a.ca, kv[1], av.y), fmaf(a.ca, kv[2], av.z), fmaf(a.ca, kv[3], av.w));
      *reinterpret_cast<float4*>(a.accu + off) = au;
      *reinterpret_cast<float4*>(a.accv + off) = av;
      ou = make_float4(fmaf(a.cs, ku[0], bu.x), fmaf(a.cs, ku[1], bu.y), fmaf(a.cs, ku[2], bu.z), fmaf(a.cs, ku[3], bu.w));
      ov = make_float4(fmaf(a.cs, kv[0], bv.x), fmaf(a.cs, kv[1], bv.y), fmaf(a.cs, kv[2], bv.z), fmaf(a.cs, kv[3], bv.w));
    }
    *reinterpret_cast<float4*>(a.outu + off) = ou;
    *reinterpret_cast<float4*>(a.outv + off) = ov;
  }
#undef SU
#undef SV
}

template <int TY, int ADV, bool SLAB, int MINB>
static int launch_explicit_mb(const StageArgs& a, const SolverParams& p, const float* forcing, cudaStream_t st) {
  constexpr int SW = TY + 8, SH = K1_TX + 2 * K1_HALO;
  const size_t smem = (size_t)2 * SH * SW * sizeof(float);
  MLSIM_SET_SMEM_ONCE((k_explicit_rk<TY, ADV, SLAB, MINB>), (int)smem);
  dim3 grid(p.ny / TY, p.nx / K1_TX + (SLAB ? 1 : 0), p.batch);
  MLSIM_CUDA_CHECK(launch_pdl(k_explicit_rk<TY, ADV, SLAB, MINB>, grid, dim3(K1_THREADS), smem, st, a, p, forcing));
  return MLSIM_OK;
}
template <int TY, int ADV, bool SLAB>
static int launch_explicit_ty(const StageArgs& a, const SolverParams& p, const float* forcing, cudaStream_t st) {
  // 3 CTAs/SM: measured against 4 (64 registers, 156 B of spills): 0.0422 vs 0.0467 ms at 256^2 x 64
  return launch_explicit_mb<TY, ADV, SLAB, 3>(a, p, forcing, st);
}

template <int ADV>
static int launch_explicit_adv(const StageArgs& a, const SolverParams& p, const float* forcing, cudaStream_t st) {
  if (p.halo > 0) {
    if (p.ny >= 128) return launch_explicit_ty<128, ADV, true>(a, p, forcing, st);
    return launch_explicit_ty<64, ADV, true>(a, p, forcing, st);
  }
  if (p.ny >= 128) return launch_explicit_ty<128, ADV, false>(a, p, forcing, st);
  return launch_explicit_ty<64, ADV, false>(a, p, forcing, st);
}

int launch_explicit(const StageArgs& a, const SolverParams& p, const float* forcing, cudaStream_t st) {
  switch (p.advection) {
    case MLSIM_ADV_UPWIND: return launch_explicit_adv<MLSIM_ADV_UPWIND>(a, p, forcing, st);
    case MLSIM_ADV_LAX_WENDROFF: return launch_explicit_adv<MLSIM_ADV_LAX_WENDROFF>(a, p, forcing, st);
    default: return launch_explicit_adv<MLSIM_ADV_VAN_LEER>(a, p, forcing, st);
  }
}

// ---------------------------------------------------------------------------------------------
// divergence (test hook; fdm.divergence of test/sim_test.py:42)
// ---------------------------------------------------------------------------------------------
__global__ void k_divergence(const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ d,
                             SolverParams p) {
  const size_t n = (size_t)p.batch * p.nx * p.ny;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % p.ny);
    const size_t row = idx / p.ny;
    const int i = (int)(row % p.nx);
    const size_t img = (row / p.nx) * (size_t)p.nx * p.ny;
    const int im = (i - 1) & (p.nx - 1), jm = (j - 1) & (p.ny - 1);
    d[idx] = (u[idx] - u[img + (size_t)im * p.ny + j]) * p.inv_hx + (v[idx] - v[img + (size_t)i * p.ny + jm]) * p.inv_hy;
  }
}

int launch_divergence(const float* u, const float* v, float* d, const SolverParams& p, cudaStream_t st) {
  k_divergence<<<148 * 8, 256, 0, st>>>(u, v, d, p);
  MLSIM_CUDA_CHECK(cudaGetLastError());
  return MLSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// Spectral buffer addressing.
//   plain:  S[b][x][nycp]                                   (x in [0, nx), all nyc columns)
//   slab :  the half spectrum is cut into `world` column chunks of kyw columns (pitch kywp) so that one
//           all-to-all moves chunk r to rank r:   S[r][b][x_local][kywp]  with `rows` = nxl (forward) or nxl+1
//           (backward: local row nxl is the x-halo of p-hat, i.e. row 0 of the next rank's slab).
// ---------------------------------------------------------------------------------------------
template <bool SLAB>
__device__ __forceinline__ size_t spec_off(const SolverParams& p, int b, int row, int k, int rows) {
  if constexpr (!SLAB) {
    return ((size_t)b * p.nx + row) * p.nycp + k;
  } else {
    const int r = k / p.kyw, kl = k - r * p.kyw;
    return (((size_t)r * p.batch + b) * rows + row) * p.kywp + kl;
  }
}

// ---------------------------------------------------------------------------------------------
// K2: divergence of (u*, v*) for 2Q rows, packed as Q complex lines (row 2q -> re, row 2q+1 -> im),
// forward FFT along y in shared memory, untangle to the two half spectra, store nyc columns per row.
// SLAB: fields are halo-extended (row -1 is the lower neighbour's last row, no wrap in x).
// ---------------------------------------------------------------------------------------------
template <int N> struct RowCfg {
  // lines per CTA such that Q*N/Rmin <= 1024 threads
  static constexpr int RMIN = FftMinRadix<N>::value;
#ifndef MLSIM_Q512
// lines per CTA for the long row transforms (measured on B200, scripts/gpu_bringup.py t=N,N,B with the variants built by
// -DMLSIM_Q..=..): 1024: Q=4 beats Q=8 (irfft_grad 0.110 -> 0.085 ms at 1024^2 x 16; 512 threads, 2 CTAs/SM instead of one
// 1024-thread CTA whose load / transform / store phases cannot overlap); 2048: Q=2 (step 2.06 -> 1.91 ms at 2048^2 x 4)
#define MLSIM_Q512 8
#define MLSIM_Q1024 4
#define MLSIM_Q2048 2
#define MLSIM_Q8192 2
#define MLSIM_C512 16
#define MLSIM_C1024 8
#define MLSIM_C2048 8
#endif
  static constexpr int Q = (N <= 256) ? 16 : (N == 512 ? MLSIM_Q512 : (N == 1024 ? MLSIM_Q1024 : (N == 2048 ? MLSIM_Q2048 : (N <= 4096 ? 2 : MLSIM_Q8192))));
  // threads per line: N / (smallest radix) runs one butterfly per thread in the narrowest stage; long lines use fewer
  // threads with two butterflies each (smaller CTAs -> more of them per SM, their phases overlap)
#ifndef MLSIM_TPLF_1024
// measured on B200 (scripts/ab_solver.py; us per launch at 1024^2 x 16 / 8192^2 x 1): K2 with 64 instead of 128 threads
// per 1024-point line 70 -> 54, with 512 instead of 1024 per 8192-point line 437 -> 352; K4 gets SLOWER with fewer
// threads (84 -> 112, 1039 -> 1196: its gradient phase is pure I/O) and keeps one butterfly per thread; the 2048-point
// column transform with 128 instead of 256 threads per column takes the 8192^2 x-solve from 637 to 511; 64 threads per
// column and 8 instead of 4 columns per CTA (64-byte row segments) 513 -> 454.
// Tried and dropped: every K2 / K4 CTA issuing bulk L2 prefetches (cp.async.bulk.prefetch.L2) of the rows the CTA one
// GPU-wide wave later will load -- K2 304 -> 334, K4 490 -> 550 us at 8192^2, 60 -> 63 / 85 -> 94 us at 1024^2 x 16.
// 8192-point lines: TWO lines per CTA (K4 then updates three of the four rows it transforms instead of one of two) with
// 512 threads per line: K4 1038 -> 489 us, K2 (256 threads per line) 352 -> 302 us, solver step 9.37 -> 7.02 ms at 8192^2
#define MLSIM_TPLF_1024 64
#define MLSIM_TPLF_2048 128
#define MLSIM_TPLF_4096 256
#define MLSIM_TPLF_8192 256
#define MLSIM_TPLI_1024 128
#define MLSIM_TPLI_2048 256
#define MLSIM_TPLI_4096 256
#define MLSIM_TPLI_8192 512
#define MLSIM_TPL_COL_1024 32
#define MLSIM_TPL_COL_2048 64
#endif
#ifndef MLSIM_K2_UNROLL
#define MLSIM_K2_UNROLL 8  // K2: columns of the divergence phase whose loads are in flight together
#endif
#ifndef MLSIM_K4_UB
#define MLSIM_K4_UB 4      // K4: columns whose loads are issued together in the gradient phase / the spectrum load
#define MLSIM_K4_UL 4
#endif
  // forward (K2) and inverse (K4) row kernels are tuned separately
  static constexpr int TPLF = (N == 1024) ? MLSIM_TPLF_1024 : (N == 2048) ? MLSIM_TPLF_2048 : (N == 4096) ? MLSIM_TPLF_4096
                            : (N == 8192) ? MLSIM_TPLF_8192 : N / RMIN;
  static constexpr int TPLI = (N == 1024) ? MLSIM_TPLI_1024 : (N == 2048) ? MLSIM_TPLI_2048 : (N == 4096) ? MLSIM_TPLI_4096
                            : (N == 8192) ? MLSIM_TPLI_8192 : N / RMIN;
  static constexpr int THREADS_F = Q * TPLF, THREADS_I = Q * TPLI;
  static constexpr size_t SMEM = fft_buf_elems<N, Q>() * sizeof(float2);
};

// Spectral column walker for the I/O phases: a thread owns ONE line and visits k = t, t + TPL, ...; in slab mode the
// chunk index / local column are advanced incrementally (no division per element) and the chunk's base pointer comes
// from the SpecPeers table (local buffer or a peer's memory).
template <bool SLAB>
struct SpecWalk {
  float2* cur;      // plain: &S[row][k]; slab: &chunk_r[b][row][0]
  size_t rowoff;    // slab: element offset of (b, row, 0) inside a chunk
  int kl, kyw, r;
  __device__ __forceinline__ void init(const SolverParams& p, float2* S, const SpecPeers& peers, int b, int row, int k,
                                       int rows) {
    if constexpr (!SLAB) {
      cur = S + ((size_t)b * p.nx + row) * p.nycp + k;
      rowoff = 0; kl = 0; kyw = 0; r = 0;
    } else {
      r = k / p.kyw;
      kl = k - r * p.kyw; kyw = p.kyw;
      rowoff = ((size_t)b * rows + row) * p.kywp;
      cur = spec_peer(peers, r) + rowoff;
    }
  }
  __device__ __forceinline__ float2* ptr() const { return SLAB ? cur + kl : cur; }
  __device__ __forceinline__ void advance(int dk, const SpecPeers& peers) {
    if constexpr (!SLAB) {
      cur += dk;
    } else {
      kl += dk;
      if (kl >= kyw) {
        while (kl >= kyw) { kl -= kyw; ++r; }
        cur = spec_peer(peers, r < MLSIM_MAX_SLAB_RANKS ? r : 0) + rowoff;
      }
    }
  }
};

template <int N, bool SLAB>
__global__ void __launch_bounds__(RowCfg<N>::THREADS_F)
k_div_rfft(const float* __restrict__ u, const float* __restrict__ v, float2* __restrict__ S, SpecPeers peers,
           const float2* __restrict__ tw, SolverParams p) {
  using C = RowCfg<N>;
  constexpr int Q = C::Q, NT = C::THREADS_F;
  constexpr int TPL = NT / Q;                       // threads per line in the I/O phases (k / y fastest)
  extern __shared__ float4 smem4[];
  float2* buf = reinterpret_cast<float2*>(smem4);

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * (2 * Q);
  const int ql = threadIdx.x / TPL, t = threadIdx.x % TPL;     // this thread's line (row pair) and lane within it

  // divergence: the thread owns row pair (2 ql, 2 ql + 1) -- the two halves of one float2 of the line-fastest layout --
  // and the columns t, t + TPL, ...: coalesced row loads through three hoisted row pointers, conflict-free 64-bit
  // shared stores
  {
    const int i = i0 + 2 * ql;
    const int im = SLAB ? (i - 1) : ((i - 1) & (p.nx - 1));
    const float* __restrict__ ru = u + (ptrdiff_t)b * p.img_stride + (ptrdiff_t)i * N;
    const float* __restrict__ rv = v + (ptrdiff_t)b * p.img_stride + (ptrdiff_t)i * N;
    const float* __restrict__ rum = u + (ptrdiff_t)b * p.img_stride + (ptrdiff_t)im * N;
    constexpr int K2U = MLSIM_K2_UNROLL;
    if (p.raw) {                    // plain transform of the field itself (spectral filter path)
      for (int j = t; j < N; j += TPL) buf[fidx<Q>(j, ql)] = make_float2(ru[j], ru[N + j]);
    } else
#pragma unroll K2U
    for (int j = t; j < N; j += TPL) {
      const int jm = (j - 1) & (N - 1);
      const float um = rum[j];
      const float ua = ru[j], ub = ru[N + j];
      const float va = rv[j], vb = rv[N + j];
      const float vam = rv[jm], vbm = rv[N + jm];
      buf[fidx<Q>(j, ql)] = make_float2((ua - um) * p.inv_hx + (va - vam) * p.inv_hy,
                                        (ub - ua) * p.inv_hx + (vb - vbm) * p.inv_hy);
    }
  }
  __syncthreads();

  const int c = threadIdx.x % Q, j = threadIdx.x / Q;
  fft_lines_forward<N, Q, C::TPLF>(buf, c, j, tw);

  // untangle the two real rows of line ql and store their half spectra
  constexpr int nyc = N / 2 + 1;
  SpecWalk<SLAB> w;
  w.init(p, S, peers, b, i0 + 2 * ql, t, p.nx);
  const int pitch = SLAB ? p.kywp : p.nycp;
  for (int k = t; k < nyc; k += TPL) {
    const float2 z = buf[fidx<Q>(k, ql)];
    const float2 zm = buf[fidx<Q>((N - k) & (N - 1), ql)];
    float2* o = w.ptr();
    o[0] = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
    o[pitch] = make_float2(0.5f * (z.y + zm.y), -0.5f * (z.x - zm.x));
    w.advance(TPL, peers);
  }
}

// ---------------------------------------------------------------------------------------------
// K4: rebuild Q complex lines from 2Q rows of p-hat, inverse FFT along y, then
//     u -= d+x p, v -= d+y p  for the first 2Q-1 rows (the last row is the x-halo of p).
// ---------------------------------------------------------------------------------------------
template <int N, bool SLAB>
__global__ void __launch_bounds__(RowCfg<N>::THREADS_I)
k_irfft_grad(const float2* __restrict__ S, SpecPeers peers, const float* uin, const float* vin,   // uin may alias uout
             float* uout, float* vout, float* __restrict__ pout,                                    // (in-place update)
             const float2* __restrict__ tw, SolverParams p) {
  using C = RowCfg<N>;
  constexpr int Q = C::Q, NT = C::THREADS_I;
  constexpr int TPL = NT / Q;
  extern __shared__ float4 smem4[];
  float2* buf = reinterpret_cast<float2*>(smem4);

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * (2 * Q - 1);
  const int ql = threadIdx.x / TPL, t = threadIdx.x % TPL;

  {
    // plain: periodic wrap; slab: the spectral rows are [0, nx] (row nx = halo), rows past it are never used
    const int ia = SLAB ? min(i0 + 2 * ql, p.nx) : ((i0 + 2 * ql) & (p.nx - 1));
    const int ib = SLAB ? min(i0 + 2 * ql + 1, p.nx) : ((i0 + 2 * ql + 1) & (p.nx - 1));
    SpecWalk<SLAB> wa, wb;
    wa.init(p, const_cast<float2*>(S), peers, b, ia, t, p.nx + 1);
    wb.init(p, const_cast<float2*>(S), peers, b, ib, t, p.nx + 1);
    // columns t, t + TPL, ... < N/2 in batches of UL: all global loads of a batch are in flight before the first of
    // them is consumed (ncu, 8192^2: a quarter of the kernel's stall samples sat on the first use of A / B)
    static_assert((N / 2) % TPL == 0, "threads per line must divide N/2");
    constexpr int MAIN = (N / 2) / TPL;
    constexpr int UL = (MAIN % MLSIM_K4_UL == 0) ? MLSIM_K4_UL : ((MAIN % 2 == 0) ? 2 : 1);
    for (int it = 0; it < MAIN; it += UL) {
      float2 A[UL], B[UL];
#pragma unroll
      for (int h = 0; h < UL; ++h) {
        A[h] = __ldg(wa.ptr());
        B[h] = __ldg(wb.ptr());
        wa.advance(TPL, peers);
        wb.advance(TPL, peers);
      }
#pragma unroll
      for (int h = 0; h < UL; ++h) {
        const int k = t + (it + h) * TPL;
        if (k == 0) {
          buf[fidx<Q>(0, ql)] = make_float2(A[h].x, B[h].x);          // c2r ignores Im of DC / Nyquist
        } else {
          buf[fidx<Q>(k, ql)] = make_float2(A[h].x - B[h].y, A[h].y + B[h].x);
          buf[fidx<Q>(N - k, ql)] = make_float2(A[h].x + B[h].y, B[h].x - A[h].y);
        }
      }
    }
    if (t == 0) {                                                      // Nyquist column (the walkers stand on it now)
      const float2 A = __ldg(wa.ptr());
      const float2 B = __ldg(wb.ptr());
      buf[fidx<Q>(N / 2, ql)] = make_float2(A.x, B.x);
    }
  }
  __syncthreads();

  const int c = threadIdx.x % Q, j = threadIdx.x / Q;
  fft_lines_inverse<N, Q, C::TPLI>(buf, c, j, tw);

  // Gradient subtraction.  A thread owns a block of 4 consecutive rows and the columns t', t' + TPG, ...: in the
  // line-fastest layout the rows 2q, 2q+1 of one y are the two halves of one float2, so its pressure values are 3 + 2
  // conflict-free 64-bit shared loads (consecutive lanes <-> consecutive y: pitch Q+1 float2), and its global accesses
  // are coalesced rows through hoisted row pointers.  All loads of an item are issued before its first store.
  constexpr int NROWS = (2 * Q - 1 > 0) ? 2 * Q - 1 : 1;
  constexpr int NRG = (NROWS + 3) / 4;                      // row groups of 4
  constexpr int TPG = (NT / NRG > N) ? N : NT / NRG;        // threads per row group (a power of two: NT, NRG are)
  const int rg = threadIdx.x / TPG, ty = threadIdx.x % TPG;
  if (rg < NRG) {
    const int r0 = 4 * rg, q0 = 2 * rg;
    const ptrdiff_t row0 = (ptrdiff_t)b * p.img_stride + (ptrdiff_t)(i0 + r0) * N;
    const float* pu = uin + row0;
    const float* pv = vin + row0;
    float* ou = uout + row0;
    float* ov = vout + row0;
    float* __restrict__ pp = pout ? pout + ((size_t)b * p.nx + i0 + r0) * N : nullptr;
    bool on[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) on[k] = (r0 + k < NROWS) && (i0 + r0 + k < p.nx);
    if (p.raw) {                    // inverse transform only: write the rows to pout
      for (int y = ty; y < N; y += TPG) {
        const float2 A0 = buf[fidx<Q>(y, q0)];
        const float2 A1 = (q0 + 1 < Q) ? buf[fidx<Q>(y, q0 + 1)] : make_float2(0.f, 0.f);
        const float pr[4] = {A0.x, A0.y, A1.x, A1.y};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (on[k]) pp[k * N + y] = pr[k];
      }
    } else
    {
      // UB columns per pass: the (possibly in-place) stores of one column would otherwise fence the next column's loads
      // behind them -- one exposed DRAM latency per column (ncu, 8192^2: 30 % of the stall samples on the first store)
      constexpr int ITERS = N / TPG;
      constexpr int UB = (ITERS % MLSIM_K4_UB == 0) ? MLSIM_K4_UB : ((ITERS % 2 == 0) ? 2 : 1);
      for (int y0 = ty; y0 < N; y0 += UB * TPG) {
        float uo[UB][4], vo[UB][4];
#pragma unroll
        for (int h = 0; h < UB; ++h) {
          const int y = y0 + h * TPG;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (on[k]) {
              uo[h][k] = pu[k * N + y];
              vo[h][k] = pv[k * N + y];
            }
          }
        }
#pragma unroll
        for (int h = 0; h < UB; ++h) {
          const int y = y0 + h * TPG;
          const int yn = (y + 1) & (N - 1);
          const float2 A0 = buf[fidx<Q>(y, q0)];
          const float2 A1 = (q0 + 1 < Q) ? buf[fidx<Q>(y, q0 + 1)] : make_float2(0.f, 0.f);
          const float2 A2 = (q0 + 2 < Q) ? buf[fidx<Q>(y, q0 + 2)] : make_float2(0.f, 0.f);
          const float2 B0 = buf[fidx<Q>(yn, q0)];
          const float2 B1 = (q0 + 1 < Q) ? buf[fidx<Q>(yn, q0 + 1)] : make_float2(0.f, 0.f);
          const float pr[5] = {A0.x, A0.y, A1.x, A1.y, A2.x};
          const float py[4] = {B0.x, B0.y, B1.x, B1.y};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (!on[k]) continue;
            ou[k * N + y] = uo[h][k] - (pr[k + 1] - pr[k]) * p.inv_hx;
            ov[k * N + y] = vo[h][k] - (py[k] - pr[k]) * p.inv_hy;
            if (pp) pp[k * N + y] = pr[k];
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3: for C adjacent ky columns: FFT along x, multiply by inv_eig (includes 1/(nx*ny)), iFFT along x.
// First stage loads straight from global, last stage stores straight to global, and the last forward
// stage hands its registers to the first inverse stage (same thread owns the same points).
// Works on a plain [image][N][nycp] buffer; image `img` uses table image img % eig_images (the split x-transform
// of long / distributed grids runs R0 sub-transforms per image, each with its own slice of the spectrum).
// ---------------------------------------------------------------------------------------------
template <int N> struct ColCfg {
  static constexpr int RMIN = FftColMinRadix<N>::value;
  static constexpr int C = (N <= 256) ? 16 : (N == 512 ? MLSIM_C512 : (N == 1024 ? MLSIM_C1024 : MLSIM_C2048));
  static constexpr int TPL = (N == 1024) ? MLSIM_TPL_COL_1024 : (N == 2048) ? MLSIM_TPL_COL_2048 : N / RMIN;     // threads per column
  static constexpr int THREADS = C * TPL;
  // at least two CTAs per SM (their load / transform / store phases overlap): caps the registers of the radix-32 stages
  static constexpr int MINB = (THREADS * 128 * 2 <= 65536) ? 2 : 1;
  static constexpr size_t SMEM = fft_buf_elems<N, C>() * sizeof(float2);
};

template <int N>
__global__ void __launch_bounds__(ColCfg<N>::THREADS, ColCfg<N>::MINB)
k_col_solve(float2* __restrict__ S, const float* __restrict__ inv_eig, const float2* __restrict__ tw,
            SolverParams p, int eig_images) {
  using P = FftColPlan<N>;
  using CC = ColCfg<N>;
  constexpr int C = CC::C;
  constexpr int RL = (P::R3 > 1) ? P::R3 : P::R2;        // radix of the last forward stage
  constexpr int NSL = N / RL;                            // its Ns
  static_assert(P::R4 == 1, "column transforms longer than 2048 go through the split passes");
  extern __shared__ float4 smem4[];
  float2* buf = reinterpret_cast<float2*>(smem4);

  pdl_launch_dependents();
  pdl_wait();
  const int c = threadIdx.x % C, j = threadIdx.x / C;
  const int ky = blockIdx.x * C + c;
  const bool valid = ky < p.nyc;
  float2* base = S + (size_t)blockIdx.y * N * p.nycp + ky;
  const float* __restrict__ eig = inv_eig + (size_t)(blockIdx.y % eig_images) * N * p.nyc;

  constexpr int TPL = CC::TPL;
  // forward stage 1: global -> registers -> smem
  {
    constexpr int R = P::R1;
    constexpr int NB = N / R, BPT = (NB + TPL - 1) / TPL;
#pragma unroll
    for (int bq = 0; bq < BPT; ++bq) {
      const int jj = j + bq * TPL;
      if (jj < NB) {
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
          v[r] = valid ? base[(size_t)(jj + r * NB) * p.nycp] : make_float2(0.f, 0.f);
        dft_regs<R, -1>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) buf[fidx<C>(jj * R + q, c)] = v[q];
      }
    }
    __syncthreads();
  }
  if constexpr (P::R3 > 1) stage_smem_t<N, P::R2, P::R1, -1, C, TPL>(buf, c, j, tw);

  // last forward stage -> scale -> first inverse stage, all in registers
  {
    constexpr int NBL = N / RL, BPT = (NBL + TPL - 1) / TPL;
    float2 v[BPT][RL];
#pragma unroll
    for (int bq = 0; bq < BPT; ++bq) {
      const int jj = j + bq * TPL;
      if (jj < NBL) {
#pragma unroll
        for (int r = 0; r < RL; ++r) v[bq][r] = buf[fidx<C>(jj + r * NSL, c)];
      }
    }
    __syncthreads();
#pragma unroll
    for (int bq = 0; bq < BPT; ++bq) {
      const int jj = j + bq * TPL;
      if (jj < NBL) {
        stage_twiddle<N, RL, NSL, -1>(v[bq], jj, tw);
        dft_regs<RL, -1>(v[bq]);                            // X[kx = jj + q*NSL]
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          const float sc = valid ? __ldg(&eig[(size_t)(jj + q * NSL) * p.nyc + ky]) : 0.f;
          v[bq][q].x *= sc;
          v[bq][q].y *= sc;
        }
        dft_regs<RL, +1>(v[bq]);                            // inverse stage 1: Ns = 1, no twiddle
#pragma unroll
        for (int q = 0; q < RL; ++q) buf[fidx<C>(jj * RL + q, c)] = v[bq][q];
      }
    }
    __syncthreads();
  }
  if constexpr (P::R3 > 1) stage_smem_t<N, P::R2, P::R3, +1, C, TPL>(buf, c, j, tw);

  // last inverse stage: smem -> registers -> global
  {
    constexpr int R = P::R1;
    constexpr int NS = N / R;
    constexpr int BPT = (NS + TPL - 1) / TPL;
#pragma unroll
    for (int bq = 0; bq < BPT; ++bq) {
      const int jj = j + bq * TPL;
      if (jj < NS) {
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = buf[fidx<C>(jj + r * NS, c)];
        stage_twiddle<N, R, NS, +1>(v, jj, tw);
        dft_regs<R, +1>(v);
        if (valid) {
#pragma unroll
          for (int q = 0; q < R; ++q) base[(size_t)(jj + q * NS) * p.nycp] = v[q];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Split x-transform (four-step FFT with an outer radix R0 done through global memory) for grids whose x extent is
// longer than one shared-memory column transform (nx > 2048) or distributed over ranks (slab decomposition):
//   nx = R0 * N2,  x = N2*n1 + n2,  kx = k1 + R0*k2.
//   pass A:  T[b][k1][n2][ky] = w^(n2*k1) * sum_n1 S[b][N2*n1 + n2][ky] e^{-2 pi i n1 k1 / R0},   w = e^{-2 pi i / nx}
//   K3    :  length-N2 transform over n2 of every (b, k1) image, times inv_eig[k1 + R0*k2][ky], inverse transform
//   pass C:  S'[b][N2*n1 + n2][ky] = sum_k1 conj(w)^(n2*k1) T[b][k1][n2][ky] e^{+2 pi i n1 k1 / R0}
// Both passes are elementwise over (b, n2, ky), coalesced along ky.  With R0 == 1 they are plain re-layouts
// (slab mode: gather the per-rank row chunks into T, scatter T into the backward chunks with the duplicated halo row).
// ---------------------------------------------------------------------------------------------
struct RowMap {          // element offset of (b, x, ky) = (x >> sh)*chunk_stride + b*img_stride + (x & mask)*pitch + ky
  int sh, mask, pitch;
  long long chunk_stride, img_stride;
  int dup_row;           // >= 0: rows with (x & mask) == 0 are also stored at local row dup_row of the previous chunk
  int nchunks;
  int use_peers;         // chunk c lives at peers.p[c] (chunk_stride ignored): peer-memory exchange
};
__device__ __forceinline__ size_t rowmap_off(const RowMap& m, int b, int x, int ky) {
  return (size_t)(x >> m.sh) * m.chunk_stride + (size_t)b * m.img_stride + (size_t)(x & m.mask) * m.pitch + ky;
}

template <int R0>
__global__ void __launch_bounds__(128)
k_xsplit_fwd(const float2* __restrict__ S, RowMap m, float2* __restrict__ T, const float2* __restrict__ twx, int n2len,
             int pitch_t) {
  pdl_launch_dependents();
  pdl_wait();
  const int ky = blockIdx.x * blockDim.x + threadIdx.x;
  if (ky >= pitch_t) return;
  const int n2 = blockIdx.y, b = blockIdx.z;
  float2 v[R0];
#pragma unroll
  for (int n1 = 0; n1 < R0; ++n1) v[n1] = __ldg(&S[rowmap_off(m, b, n1 * n2len + n2, ky)]);
  if constexpr (R0 > 1) {
    dft_regs<R0, -1>(v);
#pragma unroll
    for (int k1 = 1; k1 < R0; ++k1) {
      const float2 w = __ldg(&twx[n2 * k1]);
      const float2 x = v[k1];
      v[k1] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < R0; ++k1) T[(((size_t)b * R0 + k1) * n2len + n2) * pitch_t + ky] = v[k1];
}

template <int R0>
__global__ void __launch_bounds__(128)
k_xsplit_inv(const float2* __restrict__ T, float2* __restrict__ S, SpecPeers peers, RowMap m,
             const float2* __restrict__ twx, int n2len, int pitch_t) {
  pdl_launch_dependents();
  pdl_wait();
  const int ky = blockIdx.x * blockDim.x + threadIdx.x;
  if (ky >= pitch_t) return;
  const int n2 = blockIdx.y, b = blockIdx.z;
  float2 v[R0];
#pragma unroll
  for (int k1 = 0; k1 < R0; ++k1) v[k1] = __ldg(&T[(((size_t)b * R0 + k1) * n2len + n2) * pitch_t + ky]);
  if constexpr (R0 > 1) {
#pragma unroll
    for (int k1 = 1; k1 < R0; ++k1) {
      const float2 w = __ldg(&twx[n2 * k1]);               // conj(w)
      const float2 x = v[k1];
      v[k1] = make_float2(x.x * w.x + x.y * w.y, x.y * w.x - x.x * w.y);
    }
    dft_regs<R0, +1>(v);
  }
#pragma unroll
  for (int n1 = 0; n1 < R0; ++n1) {
    const int x = n1 * n2len + n2;
    if (!m.use_peers) {
      S[rowmap_off(m, b, x, ky)] = v[n1];
      if (m.dup_row >= 0 && (x & m.mask) == 0) {
        const int ch = ((x >> m.sh) + m.nchunks - 1) % m.nchunks;
        S[(size_t)ch * m.chunk_stride + (size_t)b * m.img_stride + (size_t)m.dup_row * m.pitch + ky] = v[n1];
      }
    } else {
      const int ch = x >> m.sh;
      spec_peer(peers, ch)[(size_t)b * m.img_stride + (size_t)(x & m.mask) * m.pitch + ky] = v[n1];
      if (m.dup_row >= 0 && (x & m.mask) == 0) {
        const int cp = (ch + m.nchunks - 1) % m.nchunks;
        spec_peer(peers, cp)[(size_t)b * m.img_stride + (size_t)m.dup_row * m.pitch + ky] = v[n1];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Whole projection in ONE kernel for small square grids: a thread-block cluster of CL CTAs owns one image and
// keeps its half spectrum in (distributed) shared memory, so the field crosses L2/HBM once (R 2w + W 2w per cell
// instead of 10w) and the three launches K2/K3/K4 become one.  CTA `rank` owns rows [rank*N/CL, (rank+1)*N/CL):
//   phase 1  div + row FFT (as K2)            -> S_local[row][ky]                      | cluster.sync
//   phase 2  column FFT * inv_eig * iFFT (K3) on its share of the ky columns, gathering / scattering the other
//            CTAs' rows through DSMEM (cluster.map_shared_rank)                        | cluster.sync
//   phase 3  row iFFT (as K4) -> p rows written over the same S_local row slots        | cluster.sync
//   phase 4  u -= d+x p, v -= d+y p (the x-halo row of p comes from the next CTA over DSMEM)
// ---------------------------------------------------------------------------------------------
template <int N> struct ProjCfg {
  static constexpr int Q = 16;                                        // complex lines (= 32 rows) per row pass
  static constexpr int THREADS = Q * N / FftMinRadix<N>::value;       // 128 (N=64) / 256 (N=128, 256)
  static constexpr int LS = Q + 1;
  static constexpr int NYC = N / 2 + 1;
  static constexpr size_t smem(int cl) { return ((size_t)(N / cl) * NYC + (size_t)N * LS) * sizeof(float2); }
};

template <int N, int CL>
__global__ void __launch_bounds__(ProjCfg<N>::THREADS)
k_project_cluster(float* __restrict__ u, float* __restrict__ v, float* __restrict__ pout,
                  const float* __restrict__ inv_eig, const float2* __restrict__ tw, SolverParams p) {
  namespace cg = cooperative_groups;
  using P = FftPlan<N>;
  using C = ProjCfg<N>;
  constexpr int Q = C::Q, LS = C::LS, NT = C::THREADS, NYC = C::NYC, RPC = N / CL;
  static_assert(RPC % (2 * Q) == 0, "rows per CTA must be a multiple of 32");
  extern __shared__ float4 smem4[];
  float2* S = reinterpret_cast<float2*>(smem4);            // [RPC][NYC] half spectrum of the CTA's rows, later p rows
  float2* buf = S + (size_t)RPC * NYC;                     // [N][LS] FFT work buffer (line-fastest)

  pdl_launch_dependents();
  pdl_wait();
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (CL == 1) ? 0 : (int)cluster.block_rank();
  const int b = blockIdx.x / CL;
  const int row0 = rank * RPC;
  const size_t img = (size_t)b * N * N;
  float* __restrict__ gu = u + img;
  float* __restrict__ gv = v + img;
  float2* Sr[CL];
#pragma unroll
  for (int r = 0; r < CL; ++r) Sr[r] = (CL == 1) ? S : cluster.map_shared_rank(S, r);

  const int c = threadIdx.x % Q, j = threadIdx.x / Q;

  // ---------------- phase 1: divergence + forward row FFT ----------------
  for (int pass = 0; pass < RPC / (2 * Q); ++pass) {
    const int i0 = row0 + pass * 2 * Q;
    for (int idx = threadIdx.x; idx < 2 * Q * N; idx += NT) {
      const int r = idx / N, y = idx - r * N;
      const int i = i0 + r;
      const int im = (i - 1) & (N - 1), ym = (y - 1) & (N - 1);
      const float d = (gu[(size_t)i * N + y] - gu[(size_t)im * N + y]) * p.inv_hx +
                      (gv[(size_t)i * N + y] - gv[(size_t)i * N + ym]) * p.inv_hy;
      reinterpret_cast<float*>(&buf[y * LS + (r >> 1)])[r & 1] = d;
    }
    __syncthreads();
    stage_smem<N, P::R1, 1, -1, Q>(buf, c, j, true, tw);
    stage_smem<N, P::R2, P::R1, -1, Q>(buf, c, j, true, tw);
    if constexpr (P::R3 > 1) stage_smem<N, P::R3, P::R1 * P::R2, -1, Q>(buf, c, j, true, tw);
    for (int idx = threadIdx.x; idx < Q * NYC; idx += NT) {
      const int q = idx / NYC, k = idx - q * NYC;
      const float2 z = buf[k * LS + q];
      const float2 zm = buf[((N - k) & (N - 1)) * LS + q];
      float2* row = S + (size_t)(pass * 2 * Q + 2 * q) * NYC;
      row[k] = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
      row[NYC + k] = make_float2(0.5f * (z.y + zm.y), -0.5f * (z.x - zm.x));
    }
    __syncthreads();
  }
  if (CL > 1) cluster.sync();

  // ---------------- phase 2: column transforms on this CTA's share of ky ----------------
  {
    constexpr int RL = (P::R3 > 1) ? P::R3 : P::R2;
    constexpr int NSL = N / RL;
    constexpr int CPC = (NYC + CL - 1) / CL;
    const int c_lo = rank * CPC, c_hi = min(NYC, c_lo + CPC);
    for (int g0 = c_lo; g0 < c_hi; g0 += Q) {
      const int ky = g0 + c;
      const bool valid = ky < c_hi;
      {
        constexpr int R = P::R1;
        if (j < N / R) {
          float2 w[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int row = j + r * (N / R);
            w[r] = valid ? Sr[row / RPC][(size_t)(row % RPC) * NYC + ky] : make_float2(0.f, 0.f);
          }
          dft_regs<R, -1>(w);
#pragma unroll
          for (int q = 0; q < R; ++q) buf[(j * R + q) * LS + c] = w[q];
        }
        __syncthreads();
      }
      if constexpr (P::R3 > 1) stage_smem<N, P::R2, P::R1, -1, Q>(buf, c, j, true, tw);
      {
        float2 w[RL];
        const bool on = j < N / RL;
        if (on) {
#pragma unroll
          for (int r = 0; r < RL; ++r) w[r] = buf[(j + r * NSL) * LS + c];
        }
        __syncthreads();
        if (on) {
          stage_twiddle<N, RL, NSL, -1>(w, j, tw);
          dft_regs<RL, -1>(w);
#pragma unroll
          for (int q = 0; q < RL; ++q) {
            const float sc = valid ? __ldg(&inv_eig[(size_t)(j + q * NSL) * p.nyc + ky]) : 0.f;
            w[q].x *= sc;
            w[q].y *= sc;
          }
          dft_regs<RL, +1>(w);
#pragma unroll
          for (int q = 0; q < RL; ++q) buf[(j * RL + q) * LS + c] = w[q];
        }
        __syncthreads();
      }
      if constexpr (P::R3 > 1) stage_smem<N, P::R2, P::R3, +1, Q>(buf, c, j, true, tw);
      {
        constexpr int R = P::R1;
        constexpr int NS = N / R;
        if (j < N / R) {
          float2 w[R];
#pragma unroll
          for (int r = 0; r < R; ++r) w[r] = buf[(j + r * (N / R)) * LS + c];
          stage_twiddle<N, R, NS, +1>(w, j, tw);
          dft_regs<R, +1>(w);
          if (valid) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
              const int row = j + q * NS;
              Sr[row / RPC][(size_t)(row % RPC) * NYC + ky] = w[q];
            }
          }
        }
        __syncthreads();
      }
    }
  }
  if (CL > 1) cluster.sync(); else __syncthreads();

  // ---------------- phase 3: inverse row FFT; p rows overwrite the row slots of S ----------------
  for (int pass = 0; pass < RPC / (2 * Q); ++pass) {
    for (int idx = threadIdx.x; idx < Q * NYC; idx += NT) {
      const int q = idx / NYC, k = idx - q * NYC;
      const float2* row = S + (size_t)(pass * 2 * Q + 2 * q) * NYC;
      const float2 A = row[k], B = row[NYC + k];
      if (k == 0 || k == N / 2) {
        buf[k * LS + q] = make_float2(A.x, B.x);
      } else {
        buf[k * LS + q] = make_float2(A.x - B.y, A.y + B.x);
        buf[(N - k) * LS + q] = make_float2(A.x + B.y, B.x - A.y);
      }
    }
    __syncthreads();
    if constexpr (P::R3 > 1) {
      stage_smem<N, P::R3, 1, +1, Q>(buf, c, j, true, tw);
      stage_smem<N, P::R2, P::R3, +1, Q>(buf, c, j, true, tw);
      stage_smem<N, P::R1, P::R3 * P::R2, +1, Q>(buf, c, j, true, tw);
    } else {
      stage_smem<N, P::R2, 1, +1, Q>(buf, c, j, true, tw);
      stage_smem<N, P::R1, P::R2, +1, Q>(buf, c, j, true, tw);
    }
    for (int idx = threadIdx.x; idx < 2 * Q * N; idx += NT) {
      const int r = idx / N, y = idx - r * N;
      reinterpret_cast<float*>(S + (size_t)(pass * 2 * Q + r) * NYC)[y] =
          reinterpret_cast<const float*>(&buf[y * LS + (r >> 1)])[r & 1];
    }
    __syncthreads();
  }
  if (CL > 1) cluster.sync();

  // ---------------- phase 4: subtract the pressure gradient ----------------
  {
    const float* pnext = reinterpret_cast<const float*>(Sr[(rank + 1) % CL]);      // row 0 of the next CTA = x-halo
    for (int idx = threadIdx.x; idx < RPC * N; idx += NT) {
      const int r = idx / N, y = idx - r * N;
      const float* prow = reinterpret_cast<const float*>(S + (size_t)r * NYC);
      const float p0 = prow[y];
      const float pyp = prow[(y + 1) & (N - 1)];
      const float pxp = (r + 1 < RPC) ? reinterpret_cast<const float*>(S + (size_t)(r + 1) * NYC)[y] : pnext[y];
      const size_t off = (size_t)(row0 + r) * N + y;
      gu[off] -= (pxp - p0) * p.inv_hx;
      gv[off] -= (pyp - p0) * p.inv_hy;
      if (pout) pout[img + off] = p0;
    }
  }
  if (CL > 1) cluster.sync();          // nobody exits while a neighbour may still read its shared memory
}

template <int N, int CL>
static int launch_project_cluster_t(float* u, float* v, float* pout, const float* inv_eig, const float2* tw,
                                    const SolverParams& p, cudaStream_t st) {
  using C = ProjCfg<N>;
  const size_t smem = C::smem(CL);
  MLSIM_SET_SMEM_ONCE((k_project_cluster<N, CL>), (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.batch * CL));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  MLSIM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_project_cluster<N, CL>, u, v, pout, inv_eig, tw, p));
  return MLSIM_OK;
}

bool project_cluster_supported(const SolverParams& p) {
  return p.nx == p.ny && (p.nx == 64 || p.nx == 128 || p.nx == 256);
}

int launch_project_cluster(float* u, float* v, float* pout, const float* inv_eig, const float2* tw,
                           const SolverParams& p, cudaStream_t st) {
  switch (p.nx) {
    case 64:  return launch_project_cluster_t<64, 1>(u, v, pout, inv_eig, tw, p, st);
    case 128: return launch_project_cluster_t<128, 1>(u, v, pout, inv_eig, tw, p, st);
    case 256: return launch_project_cluster_t<256, 4>(u, v, pout, inv_eig, tw, p, st);
    default: set_error("cluster projection: unsupported grid"); return MLSIM_EINVAL;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int N, bool SLAB>
static int launch_div_rfft_n(const float* u, const float* v, float2* S, const SpecPeers& peers, const float2* tw,
                             const SolverParams& p, cudaStream_t st) {
  using C = RowCfg<N>;
  MLSIM_SET_SMEM_ONCE((k_div_rfft<N, SLAB>), (int)C::SMEM);
  dim3 grid(p.nx / (2 * C::Q), p.batch);
  MLSIM_CUDA_CHECK(launch_pdl(k_div_rfft<N, SLAB>, grid, dim3(C::THREADS_F), C::SMEM, st, u, v, S, peers, tw, p));
  return MLSIM_OK;
}

template <int N, bool SLAB>
static int launch_irfft_grad_n(const float2* S, const SpecPeers& peers, const float* uin, const float* vin, float* uout,
                               float* vout, float* pout, const float2* tw, const SolverParams& p, cudaStream_t st) {
  using C = RowCfg<N>;
  MLSIM_SET_SMEM_ONCE((k_irfft_grad<N, SLAB>), (int)C::SMEM);
  const int rows = 2 * C::Q - 1;
  dim3 grid((p.nx + rows - 1) / rows, p.batch);
  MLSIM_CUDA_CHECK(launch_pdl(k_irfft_grad<N, SLAB>, grid, dim3(C::THREADS_I), C::SMEM, st, S, peers, uin, vin, uout, vout, pout, tw, p));
  return MLSIM_OK;
}

template <int N>
static int launch_col_solve_n(float2* S, const float* inv_eig, const float2* tw, const SolverParams& p, int images,
                              int eig_images, cudaStream_t st) {
  using C = ColCfg<N>;
  MLSIM_SET_SMEM_ONCE((k_col_solve<N>), (int)C::SMEM);
  dim3 grid((p.nyc + C::C - 1) / C::C, images);
  MLSIM_CUDA_CHECK(launch_pdl(k_col_solve<N>, grid, dim3(C::THREADS), C::SMEM, st, S, inv_eig, tw, p, eig_images));
  return MLSIM_OK;
}

#define MLSIM_DISPATCH_ROW(n, CALL)                   \
  switch (n) {                                        \
    case 64:   { constexpr int N_ = 64;   return CALL; }   \
    case 128:  { constexpr int N_ = 128;  return CALL; }   \
    case 256:  { constexpr int N_ = 256;  return CALL; }   \
    case 512:  { constexpr int N_ = 512;  return CALL; }   \
    case 1024: { constexpr int N_ = 1024; return CALL; }   \
    case 2048: { constexpr int N_ = 2048; return CALL; }   \
    case 4096: { constexpr int N_ = 4096; return CALL; }   \
    case 8192: { constexpr int N_ = 8192; return CALL; }   \
    default: set_error("unsupported ny (need a power of two in [64, 8192])"); return MLSIM_EINVAL; \
  }
#define MLSIM_DISPATCH_COL(n, CALL)                   \
  switch (n) {                                        \
    case 64:   { constexpr int N_ = 64;   return CALL; }   \
    case 128:  { constexpr int N_ = 128;  return CALL; }   \
    case 256:  { constexpr int N_ = 256;  return CALL; }   \
    case 512:  { constexpr int N_ = 512;  return CALL; }   \
    case 1024: { constexpr int N_ = 1024; return CALL; }   \
    case 2048: { constexpr int N_ = 2048; return CALL; }   \
    default: set_error("unsupported column transform length (need a power of two in [64, 2048])"); return MLSIM_EINVAL; \
  }

// slab mode: chunk r of the half spectrum goes to peers.p[r]; plain mode: S
static SpecPeers local_chunks(float2* base, size_t chunk_elems, int world) {
  SpecPeers q;
  for (int r = 0; r < MLSIM_MAX_SLAB_RANKS; ++r) q.p[r] = base ? base + (size_t)(r < world ? r : 0) * chunk_elems : nullptr;
  return q;
}
int launch_div_rfft(const float* u, const float* v, float2* S, const float2* tw_y, const SolverParams& p,
                    cudaStream_t st, const SpecPeers* peers, int world) {
  if (p.halo > 0) {
    const SpecPeers q = peers ? *peers : local_chunks(S, (size_t)p.batch * p.nx * p.kywp, world);
    MLSIM_DISPATCH_ROW(p.ny, (launch_div_rfft_n<N_, true>(u, v, S, q, tw_y, p, st)));
  }
  const SpecPeers none = local_chunks(nullptr, 0, 0);
  MLSIM_DISPATCH_ROW(p.ny, (launch_div_rfft_n<N_, false>(u, v, S, none, tw_y, p, st)));
}
int launch_irfft_grad(const float2* S, const float* uin, const float* vin, float* uout, float* vout, float* pout,
                      const float2* tw_y, const SolverParams& p, cudaStream_t st, int world) {
  if (p.halo > 0) {
    const SpecPeers q = local_chunks(const_cast<float2*>(S), (size_t)p.batch * (p.nx + 1) * p.kywp, world);
    MLSIM_DISPATCH_ROW(p.ny, (launch_irfft_grad_n<N_, true>(S, q, uin, vin, uout, vout, pout, tw_y, p, st)));
  }
  const SpecPeers none = local_chunks(nullptr, 0, 0);
  MLSIM_DISPATCH_ROW(p.ny, (launch_irfft_grad_n<N_, false>(S, none, uin, vin, uout, vout, pout, tw_y, p, st)));
}
// `n` rows per image, `images` images in the plain buffer S, image i scaled by table image i % eig_images
int launch_col_solve(float2* S, const float* inv_eig, const float2* tw_x, const SolverParams& p, int n, int images,
                     int eig_images, cudaStream_t st) {
  MLSIM_DISPATCH_COL(n, (launch_col_solve_n<N_>(S, inv_eig, tw_x, p, images, eig_images, st)));
}

// Row maps of the three spectral buffer layouts the split passes read / write.
static RowMap rowmap_plain(int nx, int pitch) {
  RowMap m; m.sh = 31; m.mask = 0x7fffffff; m.pitch = pitch; m.chunk_stride = 0; m.img_stride = (long long)nx * pitch;
  m.dup_row = -1; m.nchunks = 1; m.use_peers = 0; return m;
}
static RowMap rowmap_chunks(int nxl, int rows, int batch, int pitch, int world, bool dup) {
  RowMap m; m.sh = ilog2(nxl); m.mask = nxl - 1; m.pitch = pitch; m.img_stride = (long long)rows * pitch;
  m.chunk_stride = (long long)batch * rows * pitch; m.dup_row = dup ? nxl : -1; m.nchunks = world; m.use_peers = 0; return m;
}

template <int R0>
static int launch_xsplit_r(bool fwd, float2* S, const SpecPeers& peers, const RowMap& m, float2* T, const float2* twx,
                           int n2len, int pitch_t, int batch, cudaStream_t st) {
  dim3 grid((pitch_t + 127) / 128, n2len, batch);
  if (fwd) MLSIM_CUDA_CHECK(launch_pdl(k_xsplit_fwd<R0>, grid, dim3(128), 0, st, S, m, T, twx, n2len, pitch_t));
  else MLSIM_CUDA_CHECK(launch_pdl(k_xsplit_inv<R0>, grid, dim3(128), 0, st, T, S, peers, m, twx, n2len, pitch_t));
  return MLSIM_OK;
}
static int launch_xsplit(bool fwd, int r0, float2* S, const RowMap& m, float2* T, const float2* twx, int n2len,
                         int pitch_t, int batch, cudaStream_t st, const SpecPeers* peers_or_null = nullptr) {
  const SpecPeers peers = peers_or_null ? *peers_or_null : local_chunks(nullptr, 0, 0);
  switch (r0) {
    case 1: return launch_xsplit_r<1>(fwd, S, peers, m, T, twx, n2len, pitch_t, batch, st);
    case 2: return launch_xsplit_r<2>(fwd, S, peers, m, T, twx, n2len, pitch_t, batch, st);
    case 4: return launch_xsplit_r<4>(fwd, S, peers, m, T, twx, n2len, pitch_t, batch, st);
    case 8: return launch_xsplit_r<8>(fwd, S, peers, m, T, twx, n2len, pitch_t, batch, st);
    case 16: return launch_xsplit_r<16>(fwd, S, peers, m, T, twx, n2len, pitch_t, batch, st);
    default: set_error("unsupported outer radix of the split x-transform"); return MLSIM_EINVAL;
  }
}

// x-direction solve of a plain buffer S[b][nx][nycp] with nx = r0 * n2len > 2048:  S -> T -> (K3 on T) -> S
int launch_col_solve_split(float2* S, float2* T, const float* inv_eig_perm, const float2* twx_full, const float2* tw_n2,
                           const SolverParams& p, int r0, cudaStream_t st) {
  const int n2len = p.nx / r0;
  const RowMap m = rowmap_plain(p.nx, p.nycp);
  int rc;
  if ((rc = launch_xsplit(true, r0, S, m, T, twx_full, n2len, p.nycp, p.batch, st))) return rc;
  if ((rc = launch_col_solve(T, inv_eig_perm, tw_n2, p, n2len, p.batch * r0, r0, st))) return rc;
  return launch_xsplit(false, r0, S, m, T, twx_full, n2len, p.nycp, p.batch, st);
}

// Slab decomposition, this rank's share of the spectrum (kyv valid columns, pitch kywp), all nxg rows:
//   recv [world][b][nxl][kywp]  ->  T[b][r0][n2len][kywp]  -> K3 ->  send2 [world][b][nxl+1][kywp] (halo row duplicated)
int launch_col_solve_slab(const float2* recv, float2* send2, float2* T, const float* inv_eig_perm, const float2* twx_full,
                          const float2* tw_n2, const SolverParams& p, int nxg, int world, int kyv, int r0, cudaStream_t st,
                          const SpecPeers* peers) {
  const int n2len = nxg / r0;
  const RowMap min_ = rowmap_chunks(p.nx, p.nx, p.batch, p.kywp, world, false);
  RowMap mout = rowmap_chunks(p.nx, p.nx + 1, p.batch, p.kywp, world, true);
  mout.use_peers = peers ? 1 : 0;
  SolverParams q = p;
  q.nyc = kyv; q.nycp = p.kywp;
  int rc;
  if ((rc = launch_xsplit(true, r0, const_cast<float2*>(recv), min_, T, twx_full, n2len, p.kywp, p.batch, st))) return rc;
  if ((rc = launch_col_solve(T, inv_eig_perm, tw_n2, q, n2len, p.batch * r0, r0, st))) return rc;
  return launch_xsplit(false, r0, send2, mout, T, twx_full, n2len, p.kywp, p.batch, st, peers);
}

}  // namespace mlsim

// In-register power-of-two DFTs and Stockham autosort stage helpers used by the pressure-projection
// kernels.  The index formulas are the ones validated in tests/test_fft_indexing.py (numpy model).
//
// One Stockham stage of radix R over a line of N complex points, Ns = product of the previous radices:
//   thread j in [0, N/R):  k = j mod Ns
//     v[r]  = in[j + r*N/R]                      r = 0..R-1
//     v[r] *= exp(SIGN*2*pi*i * r*k / (Ns*R))
//     V     = DFT_R(v)
//     out[(j/Ns)*Ns*R + k + q*Ns] = V[q]         q = 0..R-1
// After the last stage (Ns*R == N) the output is in natural order.
#pragma once
#include <cuda_runtime.h>

namespace mlsim {

__host__ __device__ constexpr float cos16(int i) {
  constexpr float t[16] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                           0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                           -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
                           0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
  return t[i & 15];
}
__host__ __device__ constexpr float sin16(int i) { return cos16(i - 4); }

__host__ __device__ constexpr int bitrev(int i, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
  return r;
}
__host__ __device__ constexpr int clog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

// One radix-2 butterfly level (sub-transform length LEN) of the in-register DIT network.
template <int R, int SIGN, int LEN>
__device__ __forceinline__ void dft_level(float2 (&a)[R]) {
#pragma unroll
  for (int b = 0; b < R; b += LEN) {
#pragma unroll
    for (int k = 0; k < LEN / 2; ++k) {
      constexpr int UNIT = 32 / LEN;
      const int ti = k * UNIT;                  // twiddle angle 2*pi*ti/32, ti < 16
      const float2 x = a[b + k + LEN / 2];
      float2 t;
      if (ti == 0) {
        t = x;
      } else if (ti == 8) {                      // w = SIGN * i
        t = (SIGN > 0) ? make_float2(-x.y, x.x) : make_float2(x.y, -x.x);
      } else {
        // literal ternary chains: after unrolling `ti` is a constant and the chain folds to one immediate
        const float c = (ti == 1) ? 0.98078528040323043f
                      : (ti == 2) ? 0.92387953251128674f
                      : (ti == 3) ? 0.83146961230254524f
                      : (ti == 4) ? 0.70710678118654757f
                      : (ti == 5) ? 0.55557023301960229f
                      : (ti == 6) ? 0.38268343236508984f
                      : (ti == 7) ? 0.19509032201612833f
                      : (ti == 9) ? -0.19509032201612819f
                      : (ti == 10) ? -0.38268343236508973f
                      : (ti == 11) ? -0.55557023301960196f
                      : (ti == 12) ? -0.70710678118654746f
                      : (ti == 13) ? -0.83146961230254535f
                      : (ti == 14) ? -0.92387953251128674f
                      : (ti == 15) ? -0.98078528040323043f
                      : 0.0f;
        const float sa = (ti == 1) ? 0.19509032201612825f
                      : (ti == 2) ? 0.38268343236508978f
                      : (ti == 3) ? 0.55557023301960218f
                      : (ti == 4) ? 0.70710678118654746f
                      : (ti == 5) ? 0.83146961230254524f
                      : (ti == 6) ? 0.92387953251128674f
                      : (ti == 7) ? 0.98078528040323043f
                      : (ti == 9) ? 0.98078528040323043f
                      : (ti == 10) ? 0.92387953251128674f
                      : (ti == 11) ? 0.83146961230254546f
                      : (ti == 12) ? 0.70710678118654757f
                      : (ti == 13) ? 0.55557023301960218f
                      : (ti == 14) ? 0.38268343236508989f
                      : (ti == 15) ? 0.19509032201612861f
                      : 0.0f;
        const float s = (SIGN > 0) ? sa : -sa;
        t = make_float2(x.x * c - x.y * s, x.x * s + x.y * c);
      }
      const float2 y = a[b + k];
      a[b + k + LEN / 2] = make_float2(y.x - t.x, y.y - t.y);
      a[b + k] = make_float2(y.x + t.x, y.y + t.y);
    }
  }
}

// Natural-order in, natural-order out DFT of R in {2,4,8,16,32} points held in registers.
// SIGN = -1 forward, +1 inverse (unnormalised).
template <int R, int SIGN>
__device__ __forceinline__ void dft_regs(float2 (&v)[R]) {
  static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
  constexpr int BITS = clog2(R);
  float2 a[R];
#pragma unroll
  for (int i = 0; i < R; ++i) a[bitrev(i, BITS)] = v[i];
  dft_level<R, SIGN, 2>(a);
  if constexpr (R >= 4) dft_level<R, SIGN, 4>(a);
  if constexpr (R >= 8) dft_level<R, SIGN, 8>(a);
  if constexpr (R >= 16) dft_level<R, SIGN, 16>(a);
  if constexpr (R >= 32) dft_level<R, SIGN, 32>(a);
#pragma unroll
  for (int i = 0; i < R; ++i) v[i] = a[i];
}

// Multiply v[r] by the inter-stage twiddles.  tw[m] = exp(-2*pi*i*m/N), m in [0, N).
template <int N, int R, int NS, int SIGN>
__device__ __forceinline__ void stage_twiddle(float2 (&v)[R], int j, const float2* __restrict__ tw) {
  if (NS == 1) return;
  const int k = j & (NS - 1);
  constexpr int STEP = N / (NS * R);
#pragma unroll
  for (int r = 1; r < R; ++r) {
    float2 w = __ldg(&tw[r * k * STEP]);
    if (SIGN > 0) w.y = -w.y;
    const float2 x = v[r];
    v[r] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
  }
}

// Output index base for thread j in a stage with sub-transform size NS and radix R.
template <int R, int NS>
__device__ __forceinline__ int stage_out_base(int j) {
  return (j / NS) * NS * R + (j & (NS - 1));
}

// Shared-memory layout of a bundle of Q lines ("line-fastest": the Q lines' values of one point index are adjacent).
//   Q >= 4: pitch Q+1 per point index (conflict-free for the Stockham stages and for row-major I/O);
//   Q <= 2: long lines (N >= 4096): pitch Q with a skew of one point every 16 (a stride-8/16 first-stage store of a
//           half warp then hits 16 distinct 8-byte bank pairs instead of 2).
template <int Q>
__host__ __device__ __forceinline__ constexpr int fidx(int idx, int c) {
  return (Q >= 4) ? idx * (Q + 1) + c : (idx + (idx >> 4)) * Q + c;
}
template <int N, int Q>
__host__ __device__ constexpr size_t fft_buf_elems() {
  return (Q >= 4) ? (size_t)N * (Q + 1) : (size_t)(N + N / 16) * Q;
}

// One complete shared-memory -> shared-memory stage for the thread's butterfly j (if j < N/R) on line `c` of a
// Q-line bundle.  Contains the two __syncthreads() that make the in-place update safe; every thread of the block
// must call it.
template <int N, int R, int NS, int SIGN, int Q>
__device__ __forceinline__ void stage_smem(float2* buf, int c, int j, bool active, const float2* __restrict__ tw) {
  float2 v[R];
  const bool on = active && (j < N / R);
  if (on) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = buf[fidx<Q>(j + r * (N / R), c)];
  }
  __syncthreads();
  if (on) {
    stage_twiddle<N, R, NS, SIGN>(v, j, tw);
    dft_regs<R, SIGN>(v);
    const int j0 = stage_out_base<R, NS>(j);
#pragma unroll
    for (int q = 0; q < R; ++q) buf[fidx<Q>(j0 + q * NS, c)] = v[q];
  }
  __syncthreads();
}

// The same stage with an explicit number of threads per line (TPL): a stage of radix R has N/R butterflies per line, so
// each thread runs BPT = ceil((N/R)/TPL) of them -- all loads of the stage before its first store (in-place update).
// Fewer threads per line means smaller CTAs, hence more CTAs per SM whose load / transform / store phases overlap.
template <int N, int R, int NS, int SIGN, int Q, int TPL>
__device__ __forceinline__ void stage_smem_t(float2* buf, int c, int j, const float2* __restrict__ tw) {
  constexpr int NB = N / R;
  constexpr int BPT = (NB + TPL - 1) / TPL;
  constexpr bool FULL = (NB % TPL) == 0;
  float2 v[BPT][R];
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int jj = j + b * TPL;
    if (FULL || jj < NB) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[b][r] = buf[fidx<Q>(jj + r * NB, c)];
    }
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int jj = j + b * TPL;
    if (FULL || jj < NB) {
      stage_twiddle<N, R, NS, SIGN>(v[b], jj, tw);
      dft_regs<R, SIGN>(v[b]);
      const int j0 = stage_out_base<R, NS>(jj);
#pragma unroll
      for (int q = 0; q < R; ++q) buf[fidx<Q>(j0 + q * NS, c)] = v[b][q];
    }
  }
  __syncthreads();
}

// Radix plans per transform length (R3 == 1 / R4 == 1 mean fewer stages).  R1*R2*R3*R4 == N.
template <int N> struct FftPlan;
template <> struct FftPlan<64>   { static constexpr int R1 = 8,  R2 = 8,  R3 = 1,  R4 = 1; };
template <> struct FftPlan<128>  { static constexpr int R1 = 8,  R2 = 16, R3 = 1,  R4 = 1; };
template <> struct FftPlan<256>  { static constexpr int R1 = 16, R2 = 16, R3 = 1,  R4 = 1; };
template <> struct FftPlan<512>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 8,  R4 = 1; };
template <> struct FftPlan<1024> { static constexpr int R1 = 8,  R2 = 8,  R3 = 16, R4 = 1; };
template <> struct FftPlan<2048> { static constexpr int R1 = 8,  R2 = 16, R3 = 16, R4 = 1; };
template <> struct FftPlan<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 16, R4 = 1; };
template <> struct FftPlan<8192> { static constexpr int R1 = 8,  R2 = 8,  R3 = 8,  R4 = 16; };

// Radix plan of the column solve (K3).  Measured on B200 at 1024^2 x 16 (scripts/ab_solver.py): radix-32 butterflies
// (64 data registers, two shared-memory exchanges per direction instead of three) take K3 from 64 to 55 us; the same
// plan makes the row kernels slower (K4 84 -> 128 us: their I/O phases want many threads per line), and at 512 the
// (32, 16) plan loses to (8, 8, 8) (60 vs 54 us), so only the 1024-point column transform uses it.
template <int N> struct FftColPlan : FftPlan<N> {};
template <> struct FftColPlan<1024> { static constexpr int R1 = 32, R2 = 32, R3 = 1, R4 = 1; };
template <int N> struct FftColMinRadix {
  using P = FftColPlan<N>;
  static constexpr int a = P::R1 < P::R2 ? P::R1 : P::R2;
  static constexpr int value = (P::R3 > 1 && P::R3 < a) ? P::R3 : a;
};

template <int N> struct FftMinRadix {
  using P = FftPlan<N>;
  static constexpr int a = P::R1 < P::R2 ? P::R1 : P::R2;
  static constexpr int b = (P::R3 > 1 && P::R3 < a) ? P::R3 : a;
  static constexpr int value = (P::R4 > 1 && P::R4 < b) ? P::R4 : b;
};

// All forward stages (R1, R2, R3, R4) / all inverse stages (reversed radix order) of a Q-line bundle in shared memory,
// TPL threads per line (thread j of line c).
template <int N, int Q, int TPL>
__device__ __forceinline__ void fft_lines_forward(float2* buf, int c, int j, const float2* __restrict__ tw) {
  using P = FftPlan<N>;
  stage_smem_t<N, P::R1, 1, -1, Q, TPL>(buf, c, j, tw);
  stage_smem_t<N, P::R2, P::R1, -1, Q, TPL>(buf, c, j, tw);
  if constexpr (P::R3 > 1) stage_smem_t<N, P::R3, P::R1 * P::R2, -1, Q, TPL>(buf, c, j, tw);
  if constexpr (P::R4 > 1) stage_smem_t<N, P::R4, P::R1 * P::R2 * P::R3, -1, Q, TPL>(buf, c, j, tw);
}
template <int N, int Q, int TPL>
__device__ __forceinline__ void fft_lines_inverse(float2* buf, int c, int j, const float2* __restrict__ tw) {
  using P = FftPlan<N>;
  if constexpr (P::R4 > 1) {
    stage_smem_t<N, P::R4, 1, +1, Q, TPL>(buf, c, j, tw);
    stage_smem_t<N, P::R3, P::R4, +1, Q, TPL>(buf, c, j, tw);
    stage_smem_t<N, P::R2, P::R4 * P::R3, +1, Q, TPL>(buf, c, j, tw);
    stage_smem_t<N, P::R1, P::R4 * P::R3 * P::R2, +1, Q, TPL>(buf, c, j, tw);
  } else if constexpr (P::R3 > 1) {
    stage_smem_t<N, P::R3, 1, +1, Q, TPL>(buf, c, j, tw);
    stage_smem_t<N, P::R2, P::R3, +1, Q, TPL>(buf, c, j, tw);
    stage_smem_t<N, P::R1, P::R3 * P::R2, +1, Q, TPL>(buf, c, j, tw);
  } else {
    stage_smem_t<N, P::R2, 1, +1, Q, TPL>(buf, c, j, tw);
    stage_smem_t<N, P::R1, P::R2, +1, Q, TPL>(buf, c, j, tw);
  }
}

}  // namespace mlsim

// Learned-correction CNN (reference src/models/CNN.py:36-40 as built by buildModel, MLACFD.json):
// 3x3 / stride 1 / zero-padding 1 convolutions as implicit GEMMs on the 5th-gen tensor cores
// (tcgen05.mma, bf16 operands, fp32 accumulators in TMEM).
//
// Activation layout in HBM (plan-owned, bf16):  act[b][x][c8][yp][8]
//   c8 = channel/8 (8 planes), yp = y+1 in [0, ny+2) with zero halo columns yp=0 and yp=ny+1,
//   innermost 8 channels = one 16-byte granule.  A row-plane (fixed b,x,c8) is contiguous along y, so
//   * an input row tile [130 pixels][64 ch] is 8 bulk copies (UBLKCP) of 2080 contiguous bytes, and it
//     lands in shared memory directly in the UMMA "no-swizzle K-major" canonical layout
//     (pixel stride 16 B, 8-pixel group stride SBO = 128 B, channel-chunk stride LBO = 2080 B);
//   * the +-1 pixel shifts of the 3x3 stencil along y are just +-16 B on the descriptor start address;
//   * epilogue stores are 16 B per thread, 512 contiguous bytes per warp.
//
// Mid layers are "input stationary": one staged input row X feeds the three output rows X-1, X, X+1 in
// ONE MMA of N = 3*64 (B operand = [W(dx=+1); W(dx=0); W(dx=-1)] stacked along N) whose accumulator
// columns are three adjacent 64-column TMEM slots of an 8-slot ring.  Per input row: 3 (dy) x 4 (K16)
// MMAs of M=128 x N=192 x K=16; shared-memory operand traffic 10 KB per 96 tensor-clocks.
#include "common.cuh"
#include "tc05.cuh"
#include "conv_kernels.h"

namespace mlsim {

using namespace tc;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// max(x, 0) and round-to-nearest bf16 of two values in ONE instruction; `a` goes to the low half (even channel)
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// =============================================================================================
// First layer: 2 -> 64 channels.  K = 9 taps * 2 ch = 18, zero padded to 32 (two K16 MMAs).
// The im2col A tile is built by the threads from the fp32 state (fusing stack((u,v),1), the input
// scale and the fp32 -> bf16 conversion), so this kernel is HBM-bound on its 128 B/pixel output.
// =============================================================================================
constexpr int CF_THREADS = 128;
constexpr int CF_LBO_A = 128 * 16;     // chunk plane stride of the A tile
constexpr int CF_LBO_B = 64 * 16;      // chunk plane stride of the weight image
constexpr int CF_WBYTES = 4 * CF_LBO_B;

// Software pipeline per CTA (2 A buffers, 2 TMEM accumulators): while the MMAs of tile i run, the threads
// already hold tile i+1's global loads in flight; tile i's epilogue stores overlap tile i+1's MMAs.
__global__ void __launch_bounds__(CF_THREADS)
k_conv_first(ConvFirstArgs a) {
  __shared__ __align__(128) uint8_t sA[2][4 * CF_LBO_A];
  __shared__ __align__(128) uint8_t sW[CF_WBYTES];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < CF_WBYTES / 16; i += CF_THREADS)
    reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(a.wimg) + i);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 128);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                       // everything above (weights, barriers, TMEM) is independent of the previous kernel
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(64);

  const int nyt = (a.ny + 127) >> 7;
  const int ntiles = a.batch * a.nx * nyt;
  const size_t plane = (size_t)(a.ny + 2);

  // Every CTA walks a CONTIGUOUS range of tiles (row tiles of one image row after the other), so the position
  // (b, x, yt) is decoded once and then advanced incrementally -- no integer division on the per-tile path.
  struct Pos { int b, x, yt; };
  const int per = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = (int)blockIdx.x * per;
  const int t1 = min(t0 + per, ntiles);
  const int dir = a.reverse ? -1 : 1;
  auto decode = [&](int t) {
    const int tile = a.reverse ? (ntiles - 1 - t) : t;
    Pos q;
    q.yt = tile % nyt;
    q.x = (tile / nyt) % a.nx;
    q.b = tile / (nyt * a.nx);
    return q;
  };
  auto advance = [&](Pos q) {
    q.yt += dir;
    if (q.yt == nyt) { q.yt = 0; if (++q.x == a.nx) { q.x = 0; ++q.b; } }
    else if (q.yt < 0) { q.yt = nyt - 1; if (--q.x < 0) { q.x = a.nx - 1; --q.b; } }
    return q;
  };

  // ---- im2col row of this thread's pixel: k = (i*3 + j)*2 + ch, i <-> dx+1, j <-> dy+1 ----
  auto load_vals = [&](const Pos& q, float (&val)[18]) {
    const int y = q.yt * 128 + tid;
    const float* __restrict__ gu = a.u + (size_t)q.b * a.img_stride + (ptrdiff_t)q.x * a.ny + y;
    const float* __restrict__ gv = a.v + (size_t)q.b * a.img_stride + (ptrdiff_t)q.x * a.ny + y;
    const bool rowok[3] = {q.x > 0, true, q.x + 1 < a.nx};
    const bool colok[3] = {y > 0 && y - 1 < a.ny, y < a.ny, y + 1 < a.ny};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const bool in = rowok[i] && colok[j];
        const ptrdiff_t off = (ptrdiff_t)(i - 1) * a.ny + (j - 1);
        val[(i * 3 + j) * 2 + 0] = in ? __ldg(gu + off) : 0.0f;
        val[(i * 3 + j) * 2 + 1] = in ? __ldg(gv + off) : 0.0f;
      }
    }
  };
  auto store_a = [&](int buf, const float (&val)[18]) {
    float s[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) s[k] = val[k] * a.scale;
    uint4 c0, c1, c2;
    c0.x = pack_bf16(s[0], s[1]);   c0.y = pack_bf16(s[2], s[3]);
    c0.z = pack_bf16(s[4], s[5]);   c0.w = pack_bf16(s[6], s[7]);
    c1.x = pack_bf16(s[8], s[9]);   c1.y = pack_bf16(s[10], s[11]);
    c1.z = pack_bf16(s[12], s[13]); c1.w = pack_bf16(s[14], s[15]);
    c2.x = pack_bf16(s[16], s[17]);
    c2.y = 0x3f803f80u;                 // k = 18, 19: constant 1.0 -> the GEMM adds bias_hi + bias_lo (packed weights)
    c2.z = 0u; c2.w = 0u;
    uint8_t* A = sA[buf];
    *reinterpret_cast<uint4*>(A + 0 * CF_LBO_A + tid * 16) = c0;
    *reinterpret_cast<uint4*>(A + 1 * CF_LBO_A + tid * 16) = c1;
    *reinterpret_cast<uint4*>(A + 2 * CF_LBO_A + tid * 16) = c2;
    *reinterpret_cast<uint4*>(A + 3 * CF_LBO_A + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core (async proxy)
  };
  auto issue = [&](int buf) {
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k16 = 0; k16 < 2; ++k16) {
        const uint64_t da = umma_desc(smem_u32(sA[buf]) + k16 * 2 * CF_LBO_A, CF_LBO_A, 128);
        const uint64_t db = umma_desc(smem_u32(sW) + k16 * 2 * CF_LBO_B, CF_LBO_B, 128);
        umma_bf16(tmem + buf * 64, da, db, idesc, k16 > 0);
      }
      umma_commit(&mbar[buf]);
    }
  };

  float val[18];
  Pos cur = decode(t0 < ntiles ? t0 : 0);
  if (t0 < t1) {
    load_vals(cur, val);
    store_a(0, val);
    __syncthreads();
    issue(0);
  }
  uint32_t parity0 = 0, parity1 = 0;
  for (int t = t0, it = 0; t < t1; ++t, ++it) {
    const int buf = it & 1;
    const bool has_next = t + 1 < t1;
    const Pos nxt = advance(cur);
    if (has_next) load_vals(nxt, val);                  // loads fly while we wait for this tile's MMAs
    if (buf == 0) { mbar_wait(&mbar[0], parity0); parity0 ^= 1; }
    else          { mbar_wait(&mbar[1], parity1); parity1 ^= 1; }
    tc_fence_after();
    uint32_t r0[32], r1[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + buf * 64;
    tmem_ld32(taddr, r0);
    tmem_ld32(taddr + 32, r1);
    tmem_ld_wait();
    if (has_next) store_a(buf ^ 1, val);
    tc_fence_before();
    __syncthreads();                                    // A[buf^1] complete; accumulator `buf` fully read
    if (has_next) issue(buf ^ 1);

    const int y = cur.yt * 128 + tid;
    if (y < a.ny) {
      uint8_t* orow = reinterpret_cast<uint8_t*>(a.out) +
                      ((((size_t)cur.b * a.nx + cur.x) * 8) * plane + (size_t)(y + 1)) * 16;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        // the bias is already in the accumulator: ReLU + bf16 pack is one cvt.rn.relu per channel pair
        const uint32_t* r = (c < 4) ? &r0[c * 8] : &r1[(c - 4) * 8];
        uint4 o;
        o.x = pack_bf16_relu(__uint_as_float(r[0]), __uint_as_float(r[1]));
        o.y = pack_bf16_relu(__uint_as_float(r[2]), __uint_as_float(r[3]));
        o.z = pack_bf16_relu(__uint_as_float(r[4]), __uint_as_float(r[5]));
        o.w = pack_bf16_relu(__uint_as_float(r[6]), __uint_as_float(r[7]));
        *reinterpret_cast<uint4*>(orow + (size_t)c * plane * 16) = o;
      }
    }
    cur = nxt;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 128);
  }
}

// =============================================================================================
// Mid / last layers: 64 -> NB channels per dx block (NB = 64 mid, 16 last).
// Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2.. = epilogue (mid layers: 8 warps,
// two per TMEM lane quarter, each draining one 32-column half of the slot; last layer: 4 warps), last warp = gate
// (accumulator-slot bookkeeping, see below).
// =============================================================================================
constexpr int CM_THREADS_MID = 64 + 8 * 32 + 32;  // producer + MMA + 8 epilogue warps + gate
constexpr int CM_THREADS_LAST = 64 + 4 * 32 + 32;
#ifndef MLSIM_CM_STAGES_MID
#define MLSIM_CM_STAGES_MID 8
#endif
constexpr int CM_STAGES_MID = MLSIM_CM_STAGES_MID;   // smem input-row ring depth, mid layers (1 CTA/SM): 8 x 16.6 KB + 74 KB of weights
constexpr int CM_STAGES_LAST = 5;                // last layer: 2 CTAs/SM (issue-bound N=48 MMAs), 101.6 KB each
constexpr int CM_NPX = 130;
constexpr int CM_PLANE = CM_NPX * 16;            // 2080 B: LBO of the A operand
constexpr int CM_STAGE_BYTES = 8 * CM_PLANE;     // 16640 B
constexpr int CM_SLOTS = 8;                      // TMEM accumulator slots of NB columns
constexpr int CM_RING = 6;                       // rows cycle with period 6; slots 6,7 mirror ring positions 0,1

template <int NB> struct ConvMidSmem {
  static constexpr int CM_STAGES = (NB == 64) ? CM_STAGES_MID : CM_STAGES_LAST;
  static constexpr int LBO_B = 3 * NB * 16;
  static constexpr int WDY = 8 * LBO_B;
  static constexpr int WBYTES = 3 * WDY;
  static constexpr int BIAS_OFF = WBYTES;
  static constexpr int EXTRA_OFF = WBYTES + NB * 4;                   // 1x1 shortcut weights (fp32)
  static constexpr int EXTRA_BYTES = (NB == 64) ? 64 * 2 * 4 : 16 * 64 * 4;
  static constexpr int IMG_BYTES = EXTRA_OFF + EXTRA_BYTES;           // what the host packs and the CTA bulk-copies
  static constexpr int STAGE_OFF = ((IMG_BYTES + 127) / 128) * 128;
  static constexpr int BAR_OFF = STAGE_OFF + CM_STAGES * CM_STAGE_BYTES;
  static constexpr int NBARS = 2 * CM_STAGES + 2 * CM_SLOTS + 1;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
};

template <int NB, bool LAST>
__global__ void __launch_bounds__(LAST ? CM_THREADS_LAST : CM_THREADS_MID, LAST ? 2 : 1)
k_conv_mid(ConvMidArgs a) {
  constexpr int EPI_WARPS = LAST ? 4 : 8;
  using L = ConvMidSmem<NB>;
  constexpr int CM_STAGES = L::CM_STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  const float* sBias = reinterpret_cast<const float*>(smem + L::BIAS_OFF);
  const float* sExtra = reinterpret_cast<const float*>(smem + L::EXTRA_OFF);
  uint8_t* sStage = smem + L::STAGE_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + CM_STAGES;
  uint64_t* tfull = bars + 2 * CM_STAGES;
  uint64_t* tempty = tfull + CM_SLOTS;
  uint64_t* wbar = tempty + CM_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::NBARS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = (NB * CM_SLOTS < 32) ? 32 : NB * CM_SLOTS;

  if (tid == 0) {
    // full[s]: one arrival from the producer (with the byte count of the row's bulk copies) + one from the gate
    for (int s = 0; s < CM_STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
    for (int s = 0; s < CM_SLOTS; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EPI_WARPS); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // the weight image does not depend on the previous kernel: start its bulk copy before waiting for that kernel
  if (tid == 0) {
    mbar_arrive_expect_tx(wbar, (uint32_t)L::IMG_BYTES);
    uint32_t off = 0;
    const uint32_t total = L::IMG_BYTES;
    while (off < total) {                       // weight image + bias, in pieces of at most 32 KB
      const uint32_t n = (total - off) > 32768u ? 32768u : (total - off);
      bulk_g2s(sW + off, a.wimg + off, n, wbar);
      off += n;
    }
  }
  pdl_wait();

  const int nyt = (a.ny + 127) >> 7;
  const int nstrips = (a.nx + a.rs - 1) / a.rs;
  const int nitems = a.batch * nstrips * nyt;
  const size_t plane = (size_t)(a.ny + 2);

  // Bookkeeping shared by the roles (all incremental, no division on the per-row paths):
  //  * input rows go through the smem ring `stage` (parity flips on wrap);
  //  * output row with running index n has ring position q = n % 6 and accumulates in TMEM slot q, except that a
  //    row with q < 2 collects its first contributions in mirror slot 6+q (so that the window of input row X --
  //    rows X-1, X, X+1 at window positions 0..2 -- ALWAYS occupies the three adjacent slots q0..q0+2 <= 7,
  //    q0 = ring position of row X-1, and needs ONE N = nrows*NB segment of 12 MMAs); the epilogue adds the two;
  //  * every slot is zero when a row starts using it (the epilogue re-zeroes after draining): all MMAs accumulate.
  if (warp == 0) {
    // ------------------------------ producer (one thread) ------------------------------
    // Stages input rows as fast as shared-memory stages are released, up to CM_STAGES rows ahead of the MMA warp
    // (HBM latency under load is 2-3 rows of tensor time).
    if (lane == 0) {
      uint32_t stage = 0, empty_parity = 1;          // fresh barrier: waiting on parity 1 passes immediately
      for (int it_ = blockIdx.x; it_ < nitems; it_ += gridDim.x) {
        const int item = a.reverse ? (nitems - 1 - it_) : it_;
        const int yt = item % nyt;
        const int strip = (item / nyt) % nstrips;
        const int b = item / (nyt * nstrips);
        const int x0 = strip * a.rs;
        const int x1 = min(x0 + a.rs, a.nx);
        const int xlo = max(x0 - 1, 0), xhi = min(x1, a.nx - 1);
        const int y0 = yt * 128;
        const int npx = min(CM_NPX, a.ny + 2 - y0);
        const uint32_t bytes = (uint32_t)npx * 16u;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.in) +
                             ((((size_t)b * a.nx + xlo) * 8) * plane + (size_t)y0) * 16;
        for (int X = xlo; X <= xhi; ++X) {
          mbar_wait(&empty[stage], empty_parity);
          mbar_arrive_expect_tx(&full[stage], 8 * bytes);
          uint8_t* dst = sStage + stage * CM_STAGE_BYTES;
#pragma unroll
          for (int c = 0; c < 8; ++c) bulk_g2s(dst + c * CM_PLANE, src + (size_t)c * plane * 16, bytes, &full[stage]);
          src += (size_t)8 * plane * 16;
          if (++stage == CM_STAGES) { stage = 0; empty_parity ^= 1; }
        }
      }
    }
  } else if (warp == 2 + EPI_WARPS) {
    // ------------------------------ gate (one thread) ------------------------------
    // Owns the "accumulator slot is free" check: input row X may only be consumed once the slot(s) of the output row
    // that STARTS accumulating with it (row X+1; and row 0 of the image) have been drained and re-zeroed (tempty[q]
    // phase u = ring position q ready for its u-th use, phase 0 = initial zero fill).  It contributes the second
    // arrival of full[stage], so the MMA warp still has a single barrier to wait for per input row, while the
    // producer's loads no longer queue behind the epilogue.  (With both checks in the producer thread the loads ran at
    // most ~4 rows ahead and the MMA warp waited 298 of 1498 clk per row for its input; mlsim_conv_debug.)
    if (lane == 0) {
      uint32_t stage = 0, empty_parity = 1;
      uint32_t qn = 0, ready_par = 0;
      for (int it_ = blockIdx.x; it_ < nitems; it_ += gridDim.x) {
        const int item = a.reverse ? (nitems - 1 - it_) : it_;
        const int strip = (item / nyt) % nstrips;
        const int x0 = strip * a.rs;
        const int x1 = min(x0 + a.rs, a.nx);
        const int xlo = max(x0 - 1, 0), xhi = min(x1, a.nx - 1);
        for (int X = xlo; X <= xhi; ++X) {
          const int nnew = ((X + 1 <= x1 - 1) ? 1 : 0) + ((X == 0) ? 1 : 0);
          for (int k = 0; k < nnew; ++k) {
            mbar_wait(&tempty[qn], (ready_par >> qn) & 1u);
            ready_par ^= 1u << qn;
            qn = (qn + 1 == CM_RING) ? 0 : qn + 1;
          }
          mbar_wait(&empty[stage], empty_parity);      // the stage's previous phase is over: this arrival counts for the new one
          mbar_arrive(&full[stage]);
          if (++stage == CM_STAGES) { stage = 0; empty_parity ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (one thread) ------------------------------
    if (lane == 0) {
      mbar_wait(wbar, 0);
      tc_fence_after();
      constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
      constexpr int A_K16_STEP = (2 * CM_PLANE) / 16, B_STEP = (2 * L::LBO_B) / 16;
      static_assert(CM_RING + 2 == CM_SLOTS, "window q0..q0+2 must fit the slots");
      static_assert(8 * L::LBO_B == L::WDY, "B offsets must be uniformly spaced over (dy, k16)");
      const uint64_t db0 = ((uint64_t)DESC_HI << 32) | (uint64_t)((smem_u32(sW) >> 4) | ((uint32_t)(L::LBO_B >> 4) << 16));
      const uint64_t da0 = ((uint64_t)DESC_HI << 32) | (uint64_t)((smem_u32(sStage) >> 4) | ((uint32_t)(CM_PLANE >> 4) << 16));
      uint32_t stage = 0, full_parity = 0;
      uint32_t qrow = 0;                                // ring position of output row x0 of the current item
      long long t_full = 0, t_issue = 0, t_rows = 0;
      const long long t_begin = clock64();
      for (int it_ = blockIdx.x; it_ < nitems; it_ += gridDim.x) {
        const int item = a.reverse ? (nitems - 1 - it_) : it_;
        const int strip = (item / nyt) % nstrips;
        const int x0 = strip * a.rs;
        const int x1 = min(x0 + a.rs, a.nx);
        const int xlo = max(x0 - 1, 0), xhi = min(x1, a.nx - 1);
        // q0 = ring position of (virtual) row X-1 for the first input row of the item
        uint32_t q0 = qrow + (uint32_t)(CM_RING + xlo - 1 - x0);
        q0 = (q0 >= 2 * CM_RING) ? q0 - 2 * CM_RING : ((q0 >= CM_RING) ? q0 - CM_RING : q0);
        for (int X = xlo; X <= xhi; ++X) {
          const int lo = max(X - 1, x0), hi = min(X + 1, x1 - 1);
          const int pos_lo = lo - (X - 1);                               // 0..2: window position of row lo
          const int nrows = hi - lo + 1;
          const long long tb = clock64();
          mbar_wait(&full[stage], full_parity);
          tc_fence_after();
          const long long tc0 = clock64();
          const uint64_t da = da0 + (uint64_t)(stage * (CM_STAGE_BYTES >> 4));
          const uint64_t db = db0 + (uint64_t)(pos_lo * NB);
          umma_bf16_row12<A_K16_STEP, B_STEP>(tmem + (q0 + (uint32_t)pos_lo) * NB, da, db, umma_idesc_bf16(nrows * NB));
          umma_commit(&empty[stage]);                                  // smem stage free once the MMAs retire
          if (X - 1 >= x0) umma_commit(&tfull[q0]);                    // output row X-1 complete
          if (X == xhi && xhi == x1 - 1) umma_commit(&tfull[(q0 + 1 == CM_RING) ? 0 : q0 + 1]);   // image bottom: row X too
          t_full += tc0 - tb;
          t_issue += clock64() - tc0;
          ++t_rows;
          q0 = (q0 + 1 == CM_RING) ? 0 : q0 + 1;
          if (++stage == CM_STAGES) { stage = 0; full_parity ^= 1; }
        }
        qrow += (uint32_t)(x1 - x0);
        while (qrow >= CM_RING) qrow -= CM_RING;
      }
      if (a.dbg != nullptr) {
        unsigned long long* d = a.dbg + (size_t)blockIdx.x * 8;
        d[0] = (unsigned long long)(clock64() - t_begin);
        d[1] = 0ull;
        d[2] = (unsigned long long)t_full;
        d[3] = (unsigned long long)t_issue;
        d[4] = (unsigned long long)t_rows;
      }
    }
  } else {
    // ------------------------------ epilogue (EPI_WARPS warps over the 128 TMEM lanes) ------------------------------
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
    const int half = LAST ? 0 : ((warp - 2) >> 2);    // mid layers: which 32-column half of the 64-column slot
    // zero every accumulator slot once; afterwards each slot is re-zeroed right after it is drained
    {
      const uint32_t t0 = tmem + ((uint32_t)(quarter * 32) << 16);
      if constexpr (NB == 64) {
#pragma unroll
        for (int c = 0; c < 512; c += 64) tmem_st32_fill(t0 + c + half * 32, 0u);
      } else {
#pragma unroll
        for (int c = 0; c < NB * CM_SLOTS; c += 16) tmem_st16_fill(t0 + c, 0u);
      }
      tmem_st_wait();
      tc_fence_before();
      if (lane == 0) for (int sl = 0; sl < CM_SLOTS; ++sl) mbar_arrive(&tempty[sl]);
    }
    mbar_wait(wbar, 0);                               // bias lives behind the weight image
    uint32_t slot = 0, done_par = 0;                  // ring position of the next row to drain; per-position phase parity
    long long t_epi_wait = 0;
    const long long t_epi_begin = clock64();
    for (int it_ = blockIdx.x; it_ < nitems; it_ += gridDim.x) {
        const int item = a.reverse ? (nitems - 1 - it_) : it_;
      const int yt = item % nyt;
      const int strip = (item / nyt) % nstrips;
      const int b = item / (nyt * nstrips);
      const int x0 = strip * a.rs;
      const int x1 = min(x0 + a.rs, a.nx);
      const int y = yt * 128 + quarter * 32 + lane;
      for (int x = x0; x < x1; ++x) {
        // last layer: fetch the state this row's correction is added to BEFORE waiting for the accumulator, so the
        // global-load latency hides behind the wait instead of serialising the epilogue
        float u_old = 0.0f, v_old = 0.0f;
        uint4 rres[LAST ? 1 : 4];                         // mid layers: shortcut operand, fetched before the accumulator wait
        if constexpr (LAST) {
          if (a.y_nchw == nullptr && y < a.ny) {
            const size_t off0 = (size_t)b * a.img_stride + (size_t)x * a.ny + y;
            u_old = a.u[off0];
            v_old = a.v[off0];
          }
        } else {
          if (a.res_mode == 1 && y < a.ny) {
            const uint8_t* rrow = reinterpret_cast<const uint8_t*>(a.res) +
                                  ((((size_t)b * a.nx + x) * 8 + (size_t)(half * 4)) * plane + (size_t)(y + 1)) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) rres[c] = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)c * plane * 16));
          } else if (a.res_mode == 2 && y < a.ny) {
            const size_t off0 = (size_t)b * a.img_stride + (size_t)x * a.ny + y;
            // the network input as the first layer sees it: (u, v) * scale rounded to bf16
            u_old = __bfloat162float(__float2bfloat16_rn(a.u[off0] * a.sscale));
            v_old = __bfloat162float(__float2bfloat16_rn(a.v[off0] * a.sscale));
          }
        }
        // ring position `slot` (q); q < 2 also owns mirror slot 6+q
        const long long te0 = clock64();
        mbar_wait(&tfull[slot], (done_par >> slot) & 1u);
        t_epi_wait += clock64() - te0;
        done_par ^= 1u << slot;
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + slot * NB;
        const uint32_t maddr = taddr + CM_RING * NB;         // mirror slot (meaningful when slot < 2)
        if constexpr (!LAST) {
          uint32_t r[32];
          const uint32_t hoff = (uint32_t)(half * 32);
          tmem_ld32(taddr + hoff, r);
          tmem_ld_wait();
          tmem_st32_fill(taddr + hoff, 0u);
          if (slot < 2) {                                    // add the partial sums parked in the mirror slot
            uint32_t t[32];
            tmem_ld32(maddr + hoff, t);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(t[k]));
            tmem_st32_fill(maddr + hoff, 0u);
          }
          tmem_st_wait();
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tempty[slot]);
          if (y < a.ny) {
            uint8_t* orow = reinterpret_cast<uint8_t*>(a.out) +
                            ((((size_t)b * a.nx + x) * 8 + (size_t)(half * 4)) * plane + (size_t)(y + 1)) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[c * 8 + e]) + sBias[half * 32 + c * 8 + e];
              if (a.res_mode == 1) {
                const uint32_t w[4] = {rres[c].x, rres[c].y, rres[c].z, rres[c].w};
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  f[2 * e2] += __uint_as_float(w[e2] << 16);              // bf16 -> f32: low half = even channel
                  f[2 * e2 + 1] += __uint_as_float(w[e2] & 0xffff0000u);
                }
              } else if (a.res_mode == 2) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int ch = half * 32 + c * 8 + e;
                  f[e] = fmaf(sExtra[2 * ch], u_old, fmaf(sExtra[2 * ch + 1], v_old, f[e]));
                }
              }
              uint4 o;
              if (a.relu) {
                o.x = pack_bf16_relu(f[0], f[1]); o.y = pack_bf16_relu(f[2], f[3]);
                o.z = pack_bf16_relu(f[4], f[5]); o.w = pack_bf16_relu(f[6], f[7]);
              } else {
                o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
                o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
              }
              *reinterpret_cast<uint4*>(orow + (size_t)c * plane * 16) = o;
            }
          }
        } else {
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
          tmem_st16_fill(taddr, 0u);
          if (slot < 2) {
            uint32_t t[16];
            tmem_ld16(maddr, t);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(t[k]));
            tmem_st16_fill(maddr, 0u);
          }
          tmem_st_wait();
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tempty[slot]);
          if (y < a.ny) {
            const size_t off = (size_t)b * a.img_stride + (size_t)x * a.ny + y;
            float o[16];
#pragma unroll
            for (int co = 0; co < 16; ++co) o[co] = __uint_as_float(r[co]) + sBias[co];
            if (a.res_mode == 3) {                        // 1x1 shortcut of the block input (64 channels of this pixel)
              const uint8_t* rrow = reinterpret_cast<const uint8_t*>(a.res) + ((((size_t)b * a.nx + x) * 8) * plane + (size_t)(y + 1)) * 16;
#pragma unroll 2
              for (int c = 0; c < 8; ++c) {
                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)c * plane * 16));
                const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const float x0 = __uint_as_float(w[e2] << 16), x1 = __uint_as_float(w[e2] & 0xffff0000u);
                  const int ci = c * 8 + 2 * e2;
#pragma unroll
                  for (int co = 0; co < 4; ++co)          // the shortcut feeds at most 4 output channels (u, v correction: 2)
                    if (co < a.cout) o[co] = fmaf(sExtra[co * 64 + ci], x0, fmaf(sExtra[co * 64 + ci + 1], x1, o[co]));
                }
              }
            }
            if (a.relu) {
#pragma unroll
              for (int co = 0; co < 16; ++co) o[co] = fmaxf(o[co], 0.0f);
            }
            if (a.y_nchw != nullptr) {
              const size_t img = (size_t)a.nx * a.ny;
#pragma unroll
              for (int co = 0; co < 16; ++co)
                if (co < a.cout) a.y_nchw[((size_t)b * a.cout + co) * img + (size_t)x * a.ny + y] = o[co];
            } else {
              // fused correction add (reference src/inference.py:102): v += model(v)
              a.u[off] = u_old + o[0];
              a.v[off] = v_old + o[1];
            }
          }
        }
        slot = (slot + 1 == CM_RING) ? 0 : slot + 1;
      }
    }
    if (a.dbg != nullptr && warp == 2 && lane == 0) {
      unsigned long long* d = a.dbg + (size_t)blockIdx.x * 8;
      d[5] = (unsigned long long)(clock64() - t_epi_begin);
      d[6] = (unsigned long long)t_epi_wait;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// =============================================================================================
// launchers
// =============================================================================================
int launch_conv_first(const ConvFirstArgs& a, int sm_count, cudaStream_t st) {
  const int nyt = (a.ny + 127) / 128;
  const long ntiles = (long)a.batch * a.nx * nyt;
  const long resident = (long)sm_count * 4;      // 4 CTAs/SM: 2 x 64 TMEM columns each, ~110 registers/thread
  int grid = (int)(ntiles < resident ? ntiles : resident);
  MLSIM_CUDA_CHECK(launch_pdl(k_conv_first, dim3(grid), dim3(CF_THREADS), 0, st, a));
  return MLSIM_OK;
}

int conv_pick_strip(int nx, int ny, int batch, int sm_count) {
  const int nyt = (ny + 127) / 128;
  int rs = nx < 32 ? nx : 32;
  while (rs > 4 && (long)batch * ((nx + rs - 1) / rs) * nyt < 2L * sm_count) rs >>= 1;
  return rs;
}

template <int NB, bool LAST>
static int launch_conv_mid_t(const ConvMidArgs& a, int sm_count, cudaStream_t st) {
  using L = ConvMidSmem<NB>;
  MLSIM_SET_SMEM_ONCE((k_conv_mid<NB, LAST>), L::TOTAL);
  const int nyt = (a.ny + 127) / 128;
  const int nstrips = (a.nx + a.rs - 1) / a.rs;
  const long nitems = (long)a.batch * nstrips * nyt;
  const long resident = (long)sm_count * (LAST ? 2 : 1);
  const int grid = (int)(nitems < resident ? nitems : resident);
  MLSIM_CUDA_CHECK(launch_pdl(k_conv_mid<NB, LAST>, dim3(grid), dim3(LAST ? CM_THREADS_LAST : CM_THREADS_MID), (size_t)L::TOTAL, st, a));
  return MLSIM_OK;
}

int launch_conv_mid(const ConvMidArgs& a, int sm_count, cudaStream_t st) {
  return launch_conv_mid_t<64, false>(a, sm_count, st);
}
int launch_conv_last(const ConvMidArgs& a, int sm_count, cudaStream_t st) {
  return launch_conv_mid_t<16, true>(a, sm_count, st);
}

int conv_first_wimg_bytes() { return CF_WBYTES + 64 * 4; }
int conv_mid_wimg_bytes() { return ConvMidSmem<64>::IMG_BYTES; }
int conv_last_wimg_bytes() { return ConvMidSmem<16>::IMG_BYTES; }
int conv_mid_extra_off() { return ConvMidSmem<64>::EXTRA_OFF; }
int conv_last_extra_off() { return ConvMidSmem<16>::EXTRA_OFF; }

}  // namespace mlsim

// Plan management and the C ABI of include/mlsim.h.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges are no-ops unless a profiler is attached

#include "common.cuh"
#include "conv_kernels.h"
#include "solver_kernels.h"

namespace mlsim {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

struct CnnLayer {
  int cin = 0, cout = 0, relu = 0;
  std::vector<double> w, b;   // OIHW, bias
  bool set = false;
  uint8_t* d_wimg = nullptr;
  // shortcut added before the activation (ResNet blocks): 0 none, 1 identity, 2 1x1 conv of the network input,
  // 3 1x1 conv of an activation (last layer); src = layer whose output is the block input (-1: network input)
  int res_mode = 0, res_src = -1;
  std::vector<double> sw, sb; // 1x1 weights [cout][cin_src], bias [cout]
  int in_buf = 0, out_buf = 0, res_buf = 0;   // activation buffer assignment (finalize)
};

}  // namespace mlsim

using namespace mlsim;

// Every launching entry point: the plan's device becomes current (one plan == one device; the per-device kernel
// attributes and the launches themselves follow the current device).
#define MLSIM_ENTER(pl)                                           \
  do {                                                            \
    if (!(pl)) { set_error("null plan"); return MLSIM_EINVAL; }   \
    MLSIM_CUDA_CHECK(cudaSetDevice((pl)->cfg.device));            \
  } while (0)

namespace {
struct NvtxRange {                     // named ranges around the solver stages / the network
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace

#define MLSIM_SLAB_HALO 8          /* halo rows on each side of the halo-extended field buffers */
#define MLSIM_SLAB_CNN_HALO 7      /* rows of state the 7-layer 3x3 CNN needs from each neighbour (one per layer) */
#define MLSIM_PLAIN_ONLY(pl) do { if ((pl)->slab) { set_error("this entry point needs a plain plan (slab_world == 0); use mlsim_slab_*"); return MLSIM_ESTATE; } } while (0)
#define MLSIM_SLAB_ONLY(pl) do { if (!(pl)->slab) { set_error("this entry point needs a slab plan (slab_world >= 1)"); return MLSIM_ESTATE; } } while (0)

// Length of the in-shared-memory column transforms the split x-transform of grids with nx > 2048 is built from
// (nx = r0 * MLSIM_SPLIT_N2).  Measured at 8192^2 (scripts/ab_solver.py): 8 x 1024 (the radix-32 column kernel, two CTAs
// per SM) takes the three x-solve kernels from 454 to 379 us against 4 x 2048.
#ifndef MLSIM_SPLIT_N2
#define MLSIM_SPLIT_N2 1024
#endif

struct mlsim_plan {
  mlsim_config cfg;
  SolverParams sp;
  int sm_count = 148;
  size_t field_elems = 0;
  int64_t ws_bytes = 0;
  // solver tables / workspace
  float* d_forcing = nullptr;
  float* d_inv_eig = nullptr;
  float2* d_twx = nullptr;
  float2* d_twy = nullptr;
  float* d_w1u = nullptr; float* d_w1v = nullptr;
  float* d_w2u = nullptr; float* d_w2v = nullptr;
  float* d_accu = nullptr; float* d_accv = nullptr;
  float2* d_spec = nullptr;
  float2* d_tw_n2 = nullptr;            // twiddles of the in-smem column transform (length n2len; == d_twx when r0 == 1)
  unsigned int* d_smax = nullptr;       // per-trajectory max speed^2 (mlsim_normalize_speed)
  float* d_filt = nullptr;              // caller-supplied spectral table (mlsim_spectral_multiply), same order as d_inv_eig
  float2* d_split = nullptr;            // T buffer of the split x-transform (r0 > 1 or slab mode)
  int r0 = 1, n2len = 0;                // nx(global) = r0 * n2len, n2len <= 2048
  // slab decomposition (cfg.slab_world >= 1): this rank owns global rows [rank*nxl, (rank+1)*nxl)
  bool slab = false;
  bool have_peers = false;              // peer-memory spectral exchange (mlsim_slab_set_peers)
  SpecPeers fwd_peers, bwd_peers;       // rank r's forward / backward receive buffer, offset to this rank's slot
  int nxg = 0, nxl = 0, ext_rows = 0, kyw = 0, kywp = 0, kyv = 0, ky0 = 0;
  // cnn
  std::vector<CnnLayer> layers;
  bool cnn_ready = false;
  bool fuse_first = false;              // first layer computed inside the first hidden layer's kernel (finalize decides)
  __nv_bfloat16* d_act[3] = {nullptr, nullptr, nullptr};
  int nact = 2;
  size_t act_bytes = 0;
  int rs = 32;
  bool cluster_projection = false;      // single-kernel projection (thread-block cluster per image)
  // fp32 staging of the fp64 entry point (mlsim_step_f64)
  float* d_fu = nullptr; float* d_fv = nullptr; float* d_fp = nullptr;
  // host-rollout staging
  float* d_hu = nullptr; float* d_hv = nullptr; float* d_hp = nullptr;
  cudaStream_t own_stream = nullptr, in_stream = nullptr, out_stream = nullptr;
  static constexpr int MAX_SPLIT = 8;
  cudaStream_t split_stream[MAX_SPLIT - 1] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_SPLIT - 1] = {};
  static constexpr int MAX_CHUNKS = 8;
  int host_sizes[MAX_CHUNKS] = {};      // explicit block sizes of the host pipeline (mlsim_set_host_blocks), 0 blocks = automatic
  int host_nsizes = 0;
  cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_done[MAX_CHUNKS] = {};
  // divergence guard (mlsim_finite_check_*): device flag = first bad step, mirrored into pinned host memory
  int fc_every = 0;
  long long fc_steps = 0;
  unsigned long long* d_fcflag = nullptr;
  unsigned long long* h_fcflag = nullptr;
  // captured step graphs (mlsim_config.step_graph): one entry per (u, v, p, apply_cnn, scale) the caller steps with
  struct StepGraph {
    float* u = nullptr; float* v = nullptr; float* p = nullptr;
    int apply_cnn = 0; double scale = 0.0;
    int seen = 0;                       // calls seen with this key (the first one always runs eagerly)
    bool failed = false;                // capture or instantiation refused: this key stays on eager launches
    cudaGraphExec_t one = nullptr, many = nullptr;      // 1 step / STEP_GRAPH_UNROLL steps
  };
  static constexpr int MAX_STEP_GRAPHS = 4, STEP_GRAPH_UNROLL = 8;
  StepGraph sgraph[MAX_STEP_GRAPHS];
  int sgraph_next = 0;
  long long graph_launches = 0;         // cudaGraphLaunch calls issued so far (mlsim_step_graph_launches)
  cudaStream_t cap_stream = nullptr;
  // timing scratch
  float* d_tu = nullptr; float* d_tv = nullptr;
  unsigned long long* d_dbg = nullptr;   // per-CTA cycle counters of the conv kernels under mlsim_time_kernel
  // in-step kernel profiling (mlsim_profile_*): CUDA events around every launch of one kernel class
  int prof_class = -1;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
};

namespace {

template <class T>
int dev_alloc(mlsim_plan* pl, T** p, size_t n) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    return MLSIM_ENOMEM;
  }
  pl->ws_bytes += (int64_t)(n * sizeof(T));
  return MLSIM_OK;
}

uint16_t f32_to_bf16_rne(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40);   // NaN
  const uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return (uint16_t)(x >> 16);
}

int build_tables(mlsim_plan* pl) {
  const mlsim_config& c = pl->cfg;
  const int nx = c.nx, ny = c.ny, nyc = ny / 2 + 1;      // nx = GLOBAL row count (also in slab mode)
  const double hx = c.lx / nx, hy = c.ly / ny;
  const double PI = 3.14159265358979323846;
  std::vector<float> forcing(ny);
  for (int j = 0; j < ny; ++j)
    forcing[j] = (float)(c.forcing_scale * sin(c.forcing_wavenumber * (j + 0.5) * hy) / c.density);
  // inverse spectrum of the periodic 5-point Laplacian, 1/(nx*ny) folded in.
  // cutoff 10*eps(float64): the reference solves in fp64, so only the (0,0) mode is removed.
  // Table layout [k1 (r0)][k2 (n2len)][ky - ky0 (kyv)] with kx = k1 + r0*k2: the order in which the split
  // x-transform produces the spectrum (r0 == 1: plain [kx][ky]); slab mode keeps only this rank's columns.
  const int r0 = pl->r0, n2 = pl->n2len;
  const int ky0 = pl->slab ? pl->ky0 : 0, kyv = pl->slab ? pl->kyv : nyc;
  std::vector<float> inv((size_t)nx * kyv);
  const double cutoff = 10.0 * 2.220446049250313e-16;
  for (int k1 = 0; k1 < r0; ++k1)
    for (int k2 = 0; k2 < n2; ++k2) {
      const int kx = k1 + r0 * k2;
      const double lx = (2.0 * cos(2.0 * PI * kx / nx) - 2.0) / (hx * hx);
      for (int k = 0; k < kyv; ++k) {
        const double ly = (2.0 * cos(2.0 * PI * (ky0 + k) / ny) - 2.0) / (hy * hy);
        const double lam = lx + ly;
        inv[((size_t)k1 * n2 + k2) * kyv + k] = (fabs(lam) > cutoff) ? (float)(1.0 / lam / ((double)nx * ny)) : 0.0f;
      }
    }
  std::vector<float2> twx(nx), twy(ny), twn(n2);
  for (int m = 0; m < nx; ++m) twx[m] = make_float2((float)cos(2.0 * PI * m / nx), (float)(-sin(2.0 * PI * m / nx)));
  for (int m = 0; m < ny; ++m) twy[m] = make_float2((float)cos(2.0 * PI * m / ny), (float)(-sin(2.0 * PI * m / ny)));
  for (int m = 0; m < n2; ++m) twn[m] = make_float2((float)cos(2.0 * PI * m / n2), (float)(-sin(2.0 * PI * m / n2)));

  int rc;
  if ((rc = dev_alloc(pl, &pl->d_forcing, ny))) return rc;
  if ((rc = dev_alloc(pl, &pl->d_inv_eig, inv.size()))) return rc;
  if ((rc = dev_alloc(pl, &pl->d_twx, nx))) return rc;
  if ((rc = dev_alloc(pl, &pl->d_twy, ny))) return rc;
  if ((rc = dev_alloc(pl, &pl->d_tw_n2, n2))) return rc;
  MLSIM_CUDA_CHECK(cudaMemcpy(pl->d_forcing, forcing.data(), ny * sizeof(float), cudaMemcpyHostToDevice));
  MLSIM_CUDA_CHECK(cudaMemcpy(pl->d_inv_eig, inv.data(), inv.size() * sizeof(float), cudaMemcpyHostToDevice));
  MLSIM_CUDA_CHECK(cudaMemcpy(pl->d_twx, twx.data(), nx * sizeof(float2), cudaMemcpyHostToDevice));
  MLSIM_CUDA_CHECK(cudaMemcpy(pl->d_twy, twy.data(), ny * sizeof(float2), cudaMemcpyHostToDevice));
  MLSIM_CUDA_CHECK(cudaMemcpy(pl->d_tw_n2, twn.data(), n2 * sizeof(float2), cudaMemcpyHostToDevice));
  return MLSIM_OK;
}

struct ProfScope {
  mlsim_plan* pl; cudaStream_t st; bool on; size_t idx;
  ProfScope(mlsim_plan* pl_, int cls, cudaStream_t st_) : pl(pl_), st(st_), on(false), idx(0) {
    if (pl->prof_class == cls && pl->prof_used < pl->prof_events.size()) {
      on = true; idx = pl->prof_used++;
      cudaEventRecord(pl->prof_events[idx].first, st);
    }
  }
  ~ProfScope() { if (on) cudaEventRecord(pl->prof_events[idx].second, st); }
};

int check_field(const void* p, const char* name) {
  if (p == nullptr) {
    set_error(std::string(name) + " is null");
    return MLSIM_EINVAL;
  }
  if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) {
    set_error(std::string(name) + " is not 16-byte aligned");
    return MLSIM_EINVAL;
  }
  return MLSIM_OK;
}

// A contiguous block [b0, b0+nb) of the batch.  Field pointers handed to the helpers below are already
// offset to the block; plan-owned workspaces are offset here.
struct Chunk {
  int b0, nb;
  size_t foff(const mlsim_plan* pl) const { return (size_t)b0 * pl->cfg.nx * pl->cfg.ny; }
  size_t soff(const mlsim_plan* pl) const { return (size_t)b0 * pl->cfg.nx * pl->sp.nycp; }
  SolverParams sp(const mlsim_plan* pl) const { SolverParams q = pl->sp; q.batch = nb; return q; }
};

// P(u,v) in place on (u,v): K2 -> K3 -> K4
int project(mlsim_plan* pl, float* u, float* v, float* p_or_null, cudaStream_t st, Chunk c) {
  NvtxRange nvtx("mlsim.projection");
  int rc;
  const SolverParams sp = c.sp(pl);
  if (pl->cluster_projection) {           // small square grids: the whole projection in one cluster kernel
    ProfScope ps(pl, 7, st);
    return launch_project_cluster(u, v, p_or_null, pl->d_inv_eig, pl->d_twy, sp, st);
  }
  float2* spec = pl->d_spec + c.soff(pl);
  { ProfScope ps(pl, 1, st); if ((rc = launch_div_rfft(u, v, spec, pl->d_twy, sp, st))) return rc; }
  {
    ProfScope ps(pl, 2, st);
    if (pl->r0 == 1) rc = launch_col_solve(spec, pl->d_inv_eig, pl->d_twx, sp, sp.nx, sp.batch, 1, st);
    else rc = launch_col_solve_split(spec, pl->d_split + c.soff(pl), pl->d_inv_eig, pl->d_twx, pl->d_tw_n2, sp, pl->r0, st);
    if (rc) return rc;
  }
  ProfScope ps(pl, 3, st);
  return launch_irfft_grad(spec, u, v, u, v, p_or_null, pl->d_twy, sp, st);
}

int explicit_prof(mlsim_plan* pl, const StageArgs& a, cudaStream_t st, Chunk c) {
  NvtxRange nvtx("mlsim.explicit_terms");
  ProfScope ps(pl, 0, st);
  return launch_explicit(a, c.sp(pl), pl->d_forcing, st);
}

// One classic-RK4 step with a projection after every stage, in place on (u,v).
int solver_step(mlsim_plan* pl, float* u, float* v, float* p_or_null, cudaStream_t st, Chunk c) {
  NvtxRange nvtx("mlsim.solver_step (RK4: 4 x [explicit terms, projection])");
  const float dt = (float)pl->cfg.dt;
  const size_t o = c.foff(pl);
  float *w1u = pl->d_w1u + o, *w1v = pl->d_w1v + o, *w2u = pl->d_w2u + o, *w2v = pl->d_w2v + o;
  int rc;
  StageArgs a;
  a.u0 = u; a.v0 = v; a.accu = pl->d_accu + o; a.accv = pl->d_accv + o;
  // stage 1: k1 = F(u0); acc = u0 + dt/6 k1; u* = u0 + dt/2 k1
  a.us = u; a.vs = v; a.outu = w1u; a.outv = w1v;
  a.ca = dt / 6.0f; a.cs = 0.5f * dt; a.mode = 1;
  if ((rc = explicit_prof(pl, a, st, c))) return rc;
  if ((rc = project(pl, w1u, w1v, nullptr, st, c))) return rc;
  // stage 2: k2 = F(u1); acc += dt/3 k2; u* = u0 + dt/2 k2
  a.us = w1u; a.vs = w1v; a.outu = w2u; a.outv = w2v;
  a.ca = dt / 3.0f; a.cs = 0.5f * dt; a.mode = 2;
  if ((rc = explicit_prof(pl, a, st, c))) return rc;
  if ((rc = project(pl, w2u, w2v, nullptr, st, c))) return rc;
  // stage 3: k3 = F(u2); acc += dt/3 k3; u* = u0 + dt k3
  a.us = w2u; a.vs = w2v; a.outu = w1u; a.outv = w1v;
  a.ca = dt / 3.0f; a.cs = dt; a.mode = 2;
  if ((rc = explicit_prof(pl, a, st, c))) return rc;
  if ((rc = project(pl, w1u, w1v, nullptr, st, c))) return rc;
  // stage 4: k4 = F(u3); u* = acc + dt/6 k4  -> written over the caller's state, then projected
  a.us = w1u; a.vs = w1v; a.outu = u; a.outv = v;
  a.ca = 0.0f; a.cs = dt / 6.0f; a.mode = 3;
  if ((rc = explicit_prof(pl, a, st, c))) return rc;
  return project(pl, u, v, p_or_null, st, c);
}

// The layer chain on `rows` x ny images: first conv from the fp32 state, hidden convs between the activation buffers
// (with the ResNet shortcuts fused into their epilogues), last conv + shortcut + correction add (or y_nchw).
int run_cnn_layers(mlsim_plan* pl, const float* u_in, const float* v_in, float* u, float* v, float* y_nchw, double scale,
                   int rows, int batch, long long img_stride, size_t aoff, cudaStream_t st) {
  NvtxRange nvtx("mlsim.correction_cnn (v += model(v))");
  const mlsim_config& c = pl->cfg;
  const int L = (int)pl->layers.size();
  int rc;
  __nv_bfloat16* act[3] = {pl->d_act[0] + aoff, pl->d_act[1] + aoff, pl->d_act[2] ? pl->d_act[2] + aoff : nullptr};
  // The first layer (2 -> 64) is computed inside the first hidden layer's kernel (its 537 MB of activations at config 2
  // never reach HBM) whenever nobody else needs its output and the hidden layer is a plain 64-channel one.
  const bool fuse_first = pl->fuse_first;
  if (!fuse_first) {
    ConvFirstArgs f;
    f.u = u_in; f.v = v_in; f.out = act[pl->layers[0].out_buf]; f.wimg = pl->layers[0].d_wimg; f.scale = (float)scale;
    f.nx = rows; f.ny = c.ny; f.batch = batch; f.reverse = 0; f.img_stride = img_stride;
    ProfScope ps(pl, 4, st);
    if ((rc = launch_conv_first(f, pl->sm_count, st))) return rc;
  }
  for (int l = 1; l < L; ++l) {
    const CnnLayer& ly = pl->layers[l];
    ConvMidArgs m;
    m.in = act[ly.in_buf]; m.out = (l < L - 1) ? act[ly.out_buf] : nullptr;
    m.u = u; m.v = v; m.y_nchw = y_nchw;
    m.wimg = ly.d_wimg;
    m.nx = rows; m.ny = c.ny; m.batch = batch; m.rs = pl->rs; m.img_stride = img_stride;
    m.cout = ly.cout; m.relu = ly.relu; m.dbg = nullptr;
    m.reverse = l & 1;                  // layer 0 ascending, layer 1 descending, ...
    m.res_mode = ly.res_mode; m.res = ly.res_buf >= 0 ? act[ly.res_buf] : nullptr; m.sscale = (float)scale;
    if (ly.res_mode == 2) { m.u = const_cast<float*>(u_in); m.v = const_cast<float*>(v_in); }   // read-only here
    m.fu = u_in; m.fv = v_in; m.wimg0 = pl->layers[0].d_wimg; m.fscale = (float)scale;
    if (l == 1 && fuse_first) {
      m.reverse = 0;
      ProfScope ps(pl, 8, st);
      if ((rc = launch_conv_mid_fused(m, pl->sm_count, st))) return rc;
    } else if (l < L - 1) {
      ProfScope ps(pl, 5, st);
      if ((rc = launch_conv_mid(m, pl->sm_count, st))) return rc;
    } else {
      // (strips of 16 instead of 32 rows for this layer -- better balance over its 296 CTAs, more halo rows -- were
      //  measured: 0.137 -> 0.127 ms at 256^2 x 64, 0.461 -> 0.478 ms at 1024^2 x 16; not kept)
      ProfScope ps(pl, 6, st);
      if ((rc = launch_conv_last(m, pl->sm_count, st))) return rc;
    }
  }
  return MLSIM_OK;
}

// y = model(stack((u,v),1)*scale); either y_nchw (test hook) or fused u += y0, v += y1.
int cnn_apply(mlsim_plan* pl, const float* u_in, const float* v_in, float* u, float* v, float* y_nchw, double scale,
              cudaStream_t st, Chunk ck) {
  if (!pl->cnn_ready) {
    set_error("correction network not finalised (mlsim_cnn_begin / set_layer / finalize)");
    return MLSIM_ESTATE;
  }
  const mlsim_config& c = pl->cfg;
  const size_t aoff = (size_t)ck.b0 * c.nx * 8 * (size_t)(c.ny + 2) * 8;      // bf16 elements per image block
  // (Running the network block after block of images through one activation region, hoping that a layer would read
  // its predecessor's output from the 126 MB L2, was measured at 256^2 x 64: 1.38 ms for the whole batch at once,
  // 1.53 / 1.59 / 1.67 / 2.06 ms with blocks of 32 / 9 / 8 / 4 images -- the write stream does not stay in L2.)
  return run_cnn_layers(pl, u_in, v_in, u, v, y_nchw, scale, c.nx, ck.nb, (long long)c.nx * c.ny, aoff, st);
}

// number of independent batch blocks the solver part of a step is cut into (one stream each)
int solver_blocks(const mlsim_plan* pl, int apply_cnn, int nb) {
  int ns = pl->cfg.solver_blocks > 0 ? pl->cfg.solver_blocks : (apply_cnn ? 2 : 4);
  if (pl->cfg.solver_blocks <= 0) {
    // latency-bound shapes lose to the fork / join of the streams: measured at 64^2 x 4, bare solver step 0.176 ms on four
    // streams against 0.05 ms on one.  Split only while a block keeps at least ~200k cells (256^2 x 9 in two blocks:
    // 0.150 against 0.166 ms in one).
    const long cells = (long)nb * pl->cfg.nx * pl->cfg.ny;
    const long cap = cells / 200000;
    if (ns > cap) ns = (int)(cap < 1 ? 1 : cap);
    // Four blocks are 64 launches per bare solver step.  With blocks of ~1M cells (config 2: 256^2 x 16) the kernels are
    // 5-10 us each and the launching thread becomes the limit (scripts/solver_pass_probe.py at 256^2 x 64: host enqueue
    // 0.26-0.30 ms per step for 0.29-0.31 ms of device time with four blocks; 0.10 ms for 0.28-0.29 ms with two).  Four
    // blocks only where a block keeps >= 2M cells (1024^2 x 16 and larger).
    if (ns > 2 && cells / ns < 2000000) ns = 2;
  }
  if (ns > mlsim_plan::MAX_SPLIT) ns = mlsim_plan::MAX_SPLIT;
  if (ns > nb) ns = nb;
  return ns < 1 ? 1 : ns;
}

// nsteps of {solver step; optional correction} on one block of the batch (pointers already offset to it)
int run_steps(mlsim_plan* pl, float* u, float* v, float* p_or_null, int nsteps, int apply_cnn, double scale,
              cudaStream_t st, Chunk c, bool allow_split = true) {
  int rc;
  // The batch is cut into mlsim_config.solver_blocks blocks of independent trajectories whose solver steps run on separate
  // streams: kernels of different blocks are in different phases (load / transform / store), so their latencies
  // overlap where one big launch runs its CTAs in lockstep.  Joined before the correction network.
  // Measured at config 2 (bench.py, 200 steps): solver only 0.307 / 0.287 / 0.284 ms with 1 / 2 / 4 blocks; with the
  // correction network behind every step 1.823 (1) / 1.785 (2) / 1.805 ms (4): each join before the CNN costs a little,
  // so rollouts with the network use 2 blocks and bare solver rollouts (data generation) 4 -- where a block keeps
  // >= 2M cells; below that four blocks make the launching thread the limit (solver_blocks above).
  int ns = allow_split ? solver_blocks(pl, apply_cnn, c.nb) : 1;
  for (int s = 0; s < nsteps; ++s) {
    float* pp = (s == nsteps - 1) ? p_or_null : nullptr;
    if (ns == 1) {
      if ((rc = solver_step(pl, u, v, pp, st, c))) return rc;
    } else {
      const size_t img = (size_t)pl->cfg.nx * pl->cfg.ny;
      MLSIM_CUDA_CHECK(cudaEventRecord(pl->ev_fork, st));
      for (int k = 0; k < ns; ++k) {
        const int b0 = (int)((long)c.nb * k / ns), b1 = (int)((long)c.nb * (k + 1) / ns);
        const size_t off = (size_t)b0 * img;
        cudaStream_t sk = (k == 0) ? st : pl->split_stream[k - 1];
        if (k > 0) MLSIM_CUDA_CHECK(cudaStreamWaitEvent(sk, pl->ev_fork, 0));
        if ((rc = solver_step(pl, u + off, v + off, pp ? pp + off : nullptr, sk, Chunk{c.b0 + b0, b1 - b0}))) return rc;
        if (k > 0) {
          MLSIM_CUDA_CHECK(cudaEventRecord(pl->ev_join[k - 1], sk));
          MLSIM_CUDA_CHECK(cudaStreamWaitEvent(st, pl->ev_join[k - 1], 0));
        }
      }
    }
    if (apply_cnn && (rc = cnn_apply(pl, u, v, u, v, nullptr, scale, st, c))) return rc;
    if (pl->fc_every > 0 && (pl->fc_steps + s + 1) % pl->fc_every == 0) {
      if ((rc = launch_finite_check(u, v, (size_t)c.nb * pl->cfg.nx * pl->cfg.ny, (unsigned long long)(pl->fc_steps + s + 1),
                                    pl->d_fcflag, st))) return rc;
      MLSIM_CUDA_CHECK(cudaMemcpyAsync(pl->h_fcflag, pl->d_fcflag, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
  }
  return MLSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// Step graphs (mlsim_config.step_graph = 1).  mlsim_step / mlsim_step_f64 replay a captured CUDA graph of {solver step;
// correction} instead of launching its kernels one by one: the same kernels with the same arguments in the same order
// (results bit-identical to the eager launches), programmatic-dependent-launch edges and the fork / join of the batch
// blocks included.  A graph bakes its pointers in, so graphs are kept per (u, v, p, apply_cnn, scale); the first call
// with a new key runs eagerly (it also sets the per-device kernel attributes, which must not happen inside a capture),
// the second captures.  Anything a capture refuses leaves that key on the eager path -- the same kernels, not another
// implementation.
// Opt-in, because it buys host time only: measured on B200 (scripts/graph_latency.py) the 64^2 x 1 step takes 92.4 us
// replayed against 92.9 us launched eagerly (50.4 / 50.3 without the network) -- the latency-bound shapes are bound by
// the dependent-launch latency of their 29 kernels ON THE DEVICE (~3.2 us each), not by the launching thread; what the
// replay saves is the ~0.1 ms of host enqueue time per step.
// ---------------------------------------------------------------------------------------------
void release_graph(mlsim_plan::StepGraph& g) {
  if (g.one) cudaGraphExecDestroy(g.one);
  if (g.many) cudaGraphExecDestroy(g.many);
  g = mlsim_plan::StepGraph();
}
void drop_graphs(mlsim_plan* pl) {
  bool any = false;
  for (auto& g : pl->sgraph) any = any || g.one || g.many;
  if (any) cudaDeviceSynchronize();     // a replay may still be in flight on the caller's stream
  for (auto& g : pl->sgraph) release_graph(g);
}

bool graph_eligible(const mlsim_plan* pl, int apply_cnn) {
  (void)apply_cnn;
  return pl->cfg.step_graph > 0 && !pl->slab && pl->fc_every == 0 && pl->prof_class < 0;
}

int capture_steps(mlsim_plan* pl, float* u, float* v, float* p, int n, int apply_cnn, double scale, cudaGraphExec_t* out) {
  *out = nullptr;
  if (!pl->cap_stream && cudaStreamCreateWithFlags(&pl->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    return MLSIM_ECUDA;
  }
  if (cudaStreamBeginCapture(pl->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return MLSIM_ECUDA;
  }
  const int rc = run_steps(pl, u, v, p, n, apply_cnn, scale, pl->cap_stream, Chunk{0, pl->cfg.batch});
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(pl->cap_stream, &g);      // always leave capture mode
  if (rc || e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return rc ? rc : MLSIM_ECUDA;
  }
  e = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return MLSIM_ECUDA;
  }
  return MLSIM_OK;
}

// nsteps of the whole batch: graph replay where the plan allows it, eager launches otherwise
int run_steps_graphed(mlsim_plan* pl, float* u, float* v, float* p, int nsteps, int apply_cnn, double scale, cudaStream_t st) {
  const Chunk all{0, pl->cfg.batch};
  if (nsteps < 1 || !graph_eligible(pl, apply_cnn)) return run_steps(pl, u, v, p, nsteps, apply_cnn, scale, st, all);
  mlsim_plan::StepGraph* g = nullptr;
  for (auto& e : pl->sgraph)
    if (e.seen && e.u == u && e.v == v && e.p == p && e.apply_cnn == apply_cnn && e.scale == scale) g = &e;
  if (!g) {
    g = &pl->sgraph[pl->sgraph_next];
    pl->sgraph_next = (pl->sgraph_next + 1) % mlsim_plan::MAX_STEP_GRAPHS;
    if (g->one || g->many) cudaDeviceSynchronize();       // evicting an entry whose replay may still be in flight
    release_graph(*g);
    g->u = u; g->v = v; g->p = p; g->apply_cnn = apply_cnn; g->scale = scale;
  }
  if (g->seen++ == 0 || g->failed) return run_steps(pl, u, v, p, nsteps, apply_cnn, scale, st, all);
  constexpr int UN = mlsim_plan::STEP_GRAPH_UNROLL;
  if (!g->one && capture_steps(pl, u, v, p, 1, apply_cnn, scale, &g->one)) g->failed = true;
  if (!g->failed && nsteps >= UN && !g->many && capture_steps(pl, u, v, p, UN, apply_cnn, scale, &g->many)) g->failed = true;
  if (g->failed) return run_steps(pl, u, v, p, nsteps, apply_cnn, scale, st, all);
  // (with a pressure output every replayed graph writes p in its last step where the eager loop writes it once, in the
  // last step of the call: the final content is the same)
  int left = nsteps;
  for (; left >= UN && g->many; left -= UN, ++pl->graph_launches) MLSIM_CUDA_CHECK(cudaGraphLaunch(g->many, st));
  for (; left > 0; --left, ++pl->graph_launches) MLSIM_CUDA_CHECK(cudaGraphLaunch(g->one, st));
  return MLSIM_OK;
}

}  // namespace

extern "C" {

const char* mlsim_last_error(void) { return g_last_error.c_str(); }
const char* mlsim_version(void) { return "mlsim-b200 0.1 (sm_100a)"; }

int mlsim_plan_create(const mlsim_config* cfg, mlsim_plan** out) {
  if (!cfg || !out) { set_error("null argument"); return MLSIM_EINVAL; }
  *out = nullptr;
  MLSIM_REQUIRE(is_pow2(cfg->nx) && cfg->nx >= 64 && cfg->nx <= 8192, "nx must be a power of two in [64, 8192]");
  MLSIM_REQUIRE(is_pow2(cfg->ny) && cfg->ny >= 64 && cfg->ny <= 8192, "ny must be a power of two in [64, 8192]");
  MLSIM_REQUIRE(cfg->batch >= 1 && cfg->batch <= 65535, "batch must be in [1, 65535]");
  MLSIM_REQUIRE(cfg->lx > 0 && cfg->ly > 0 && cfg->dt > 0 && cfg->density > 0, "lx, ly, dt, density must be positive");
  MLSIM_REQUIRE(cfg->advection >= 0 && cfg->advection <= 2, "unknown advection scheme");
  MLSIM_REQUIRE(cfg->solver_blocks >= 0 && cfg->solver_blocks <= mlsim_plan::MAX_SPLIT, "solver_blocks must be in [0, 8]");
  MLSIM_REQUIRE(cfg->host_blocks >= 0 && cfg->host_blocks <= mlsim_plan::MAX_CHUNKS, "host_blocks must be in [0, 8]");
  MLSIM_REQUIRE(cfg->step_graph == 0 || cfg->step_graph == 1, "step_graph must be 0 (eager launches) or 1 (graph replay)");
  // KolmogorovForcing(diam, wave_number): the reference only ever builds it with diam = domain length = 2*pi
  // (src/inference.py:61,81-82), where "sin(k y)" and "sin(2 pi k y / diam)" coincide.  For any other domain the two
  // readings of the third-party formula differ and neither can be checked here, so it is refused rather than guessed.
  MLSIM_REQUIRE(cfg->forcing_scale == 0.0 || fabs(cfg->ly - 2.0 * 3.14159265358979323846) < 1e-9,
                "Kolmogorov forcing is defined here for ly = 2*pi only (set forcing_scale = 0 for other domains)");
  MLSIM_REQUIRE((long)cfg->batch * (cfg->nx > 2048 ? cfg->nx / MLSIM_SPLIT_N2 : 1) <= 65535,
                "batch * (number of x sub-transforms) must be <= 65535");
  const bool slab = cfg->slab_world >= 1;
  if (slab) {
    const int W = cfg->slab_world;
    MLSIM_REQUIRE(is_pow2(W) && W <= MLSIM_MAX_SLAB_RANKS, "slab_world must be a power of two <= 8 (one NVLink domain)");
    MLSIM_REQUIRE(cfg->slab_rank >= 0 && cfg->slab_rank < W, "slab_rank out of range");
    MLSIM_REQUIRE(cfg->nx / W >= 32, "a slab needs at least 32 rows per rank");
    MLSIM_REQUIRE(cfg->ny / 2 >= W * W, "ny too small for this many ranks (every rank needs spectral columns)");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    return MLSIM_ECUDA;
  }
  MLSIM_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
  MLSIM_CUDA_CHECK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  MLSIM_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    set_error("this library contains sm_100a code only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    return MLSIM_ECUDA;
  }
  mlsim_plan* pl = new mlsim_plan();
  pl->cfg = *cfg;
  pl->sm_count = prop.multiProcessorCount;
  const double hx = cfg->lx / cfg->nx, hy = cfg->ly / cfg->ny;
  SolverParams& sp = pl->sp;
  const int nyc = cfg->ny / 2 + 1;
  pl->slab = slab;
  pl->nxg = cfg->nx;
  pl->r0 = cfg->nx > 2048 ? cfg->nx / MLSIM_SPLIT_N2 : 1;
  pl->n2len = cfg->nx / pl->r0;
  if (slab) {
    const int W = cfg->slab_world;
    pl->nxl = cfg->nx / W;
    pl->ext_rows = pl->nxl + 2 * MLSIM_SLAB_HALO;
    pl->kyw = (nyc + W - 1) / W;
    pl->kywp = (pl->kyw + 3) & ~3;
    pl->ky0 = cfg->slab_rank * pl->kyw;
    pl->kyv = nyc - pl->ky0 < pl->kyw ? nyc - pl->ky0 : pl->kyw;
  }
  sp.nx = slab ? pl->nxl : cfg->nx; sp.ny = cfg->ny; sp.batch = cfg->batch;
  sp.nyc = nyc;
  sp.nycp = (sp.nyc + 3) & ~3;
  sp.inv_hx = (float)(1.0 / hx); sp.inv_hy = (float)(1.0 / hy);
  sp.nu_hx2 = (float)(cfg->viscosity / cfg->density / (hx * hx));
  sp.nu_hy2 = (float)(cfg->viscosity / cfg->density / (hy * hy));
  sp.dt_hx = (float)(cfg->lw_dt_scale * cfg->dt / hx);
  sp.dt_hy = (float)(cfg->lw_dt_scale * cfg->dt / hy);
  sp.drag = (float)cfg->drag;
  sp.advection = cfg->advection;
  sp.img_stride = slab ? (long long)pl->ext_rows * cfg->ny : (long long)cfg->nx * cfg->ny;
  sp.halo = slab ? MLSIM_SLAB_HALO : 0;
  sp.kyw = pl->kyw; sp.kywp = pl->kywp;
  sp.raw = 0;
  pl->field_elems = (size_t)cfg->batch * sp.nx * cfg->ny;
  // Single-kernel projection (thread-block cluster per image, DSMEM half spectrum): measured on B200 against the
  // three-kernel path -- 256^2 x 64: 0.0665 vs 0.0525 ms per projection; after the K2/K4 rewrites also 128^2 x 512
  // (solver step 0.711 vs 0.565 ms) and 64^2 x 2048 (0.542 vs 0.529 ms).  Opt-in only: mlsim_config.cluster_projection.
  if (cfg->cluster_projection && (slab || !project_cluster_supported(sp))) {
    set_error("cluster_projection needs a plain square 64 / 128 / 256 grid");
    delete pl;
    return MLSIM_EINVAL;
  }
  pl->cluster_projection = cfg->cluster_projection != 0;

  int rc = build_tables(pl);
  const size_t n = pl->field_elems;
  if (!slab) {
    // plain mode: the plan owns the RK workspaces and the spectral buffer(s); slab mode: the caller owns every
    // exchanged buffer (halo-extended fields, spectral send/recv chunks), the plan only the split-transform buffer
    if (!rc) rc = dev_alloc(pl, &pl->d_w1u, n);
    if (!rc) rc = dev_alloc(pl, &pl->d_w1v, n);
    if (!rc) rc = dev_alloc(pl, &pl->d_w2u, n);
    if (!rc) rc = dev_alloc(pl, &pl->d_w2v, n);
    if (!rc) rc = dev_alloc(pl, &pl->d_accu, n);
    if (!rc) rc = dev_alloc(pl, &pl->d_accv, n);
    const size_t ns = (size_t)cfg->batch * cfg->nx * sp.nycp;
    if (!rc) rc = dev_alloc(pl, &pl->d_spec, ns);
    if (!rc) {
      cudaError_t e2 = cudaMemset(pl->d_spec, 0, ns * sizeof(float2));
      if (e2 != cudaSuccess) { set_error(cudaGetErrorString(e2)); rc = MLSIM_ECUDA; }
    }
    if (!rc && pl->r0 > 1) rc = dev_alloc(pl, &pl->d_split, ns);
  } else {
    const size_t ns = (size_t)cfg->batch * cfg->nx * pl->kywp;
    if (!rc) rc = dev_alloc(pl, &pl->d_split, ns);
    if (!rc) {
      cudaError_t e2 = cudaMemset(pl->d_split, 0, ns * sizeof(float2));
      if (e2 != cudaSuccess) { set_error(cudaGetErrorString(e2)); rc = MLSIM_ECUDA; }
    }
  }
  if (!rc) {
    cudaError_t e2 = cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < mlsim_plan::MAX_SPLIT - 1 && e2 == cudaSuccess; ++i) {
      e2 = cudaStreamCreateWithFlags(&pl->split_stream[i], cudaStreamNonBlocking);
      if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&pl->ev_join[i], cudaEventDisableTiming);
    }
    if (e2 != cudaSuccess) { set_error(cudaGetErrorString(e2)); rc = MLSIM_ECUDA; }
  }
  if (rc) { mlsim_plan_destroy(pl); return rc; }
  *out = pl;
  return MLSIM_OK;
}

int mlsim_plan_destroy(mlsim_plan* pl) {
  if (!pl) return MLSIM_OK;
  cudaSetDevice(pl->cfg.device);
  cudaDeviceSynchronize();
  void* ptrs[] = {pl->d_forcing, pl->d_inv_eig, pl->d_twx, pl->d_twy, pl->d_w1u, pl->d_w1v, pl->d_w2u, pl->d_w2v,
                  pl->d_accu, pl->d_accv, pl->d_spec, pl->d_tw_n2, pl->d_split, pl->d_filt, pl->d_smax, pl->d_act[0], pl->d_act[1], pl->d_act[2], pl->d_hu, pl->d_hv, pl->d_hp, pl->d_fu, pl->d_fv, pl->d_fp,
                  pl->d_tu, pl->d_tv, pl->d_dbg};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (pl->d_fcflag) cudaFree(pl->d_fcflag);
  if (pl->h_fcflag) cudaFreeHost(pl->h_fcflag);
  for (auto& l : pl->layers) if (l.d_wimg) cudaFree(l.d_wimg);
  drop_graphs(pl);
  if (pl->cap_stream) cudaStreamDestroy(pl->cap_stream);
  if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
  for (int i = 0; i < mlsim_plan::MAX_SPLIT - 1; ++i) {
    if (pl->split_stream[i]) cudaStreamDestroy(pl->split_stream[i]);
    if (pl->ev_join[i]) cudaEventDestroy(pl->ev_join[i]);
  }
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->in_stream) cudaStreamDestroy(pl->in_stream);
  if (pl->out_stream) cudaStreamDestroy(pl->out_stream);
  for (int i = 0; i < mlsim_plan::MAX_CHUNKS; ++i) { if (pl->ev_in[i]) cudaEventDestroy(pl->ev_in[i]); if (pl->ev_done[i]) cudaEventDestroy(pl->ev_done[i]); }
  for (auto& e : pl->prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  delete pl;
  return MLSIM_OK;
}


// Host-only: pack one conv layer (OIHW fp64 weights + bias) into the bf16 operand image the kernels
// bulk-copy into shared memory.  kind 0 = first layer (K = 18 padded to 32, N = 64), 1 = mid layer
// (NB = 64), 2 = last layer (NB = 16).  With out == nullptr returns the image size in bytes.
int64_t mlsim_cnn_pack_host(int32_t kind, int32_t cin, int32_t cout, const double* w, const double* bias, void* out,
                            int64_t out_bytes) {
  const int64_t need = kind == 0 ? conv_first_wimg_bytes() : (kind == 1 ? conv_mid_wimg_bytes() : conv_last_wimg_bytes());
  if (out == nullptr) return need;
  if (out_bytes < need || !w || !bias || kind < 0 || kind > 2) { set_error("mlsim_cnn_pack_host: bad arguments"); return MLSIM_EINVAL; }
  uint8_t* img = static_cast<uint8_t*>(out);
  memset(img, 0, (size_t)need);
  auto W = [&](int co, int ci, int i, int j) -> float {
    if (co >= cout || ci >= cin) return 0.0f;
    return (float)w[(((size_t)co * cin + ci) * 3 + i) * 3 + j];
  };
  uint16_t* q = reinterpret_cast<uint16_t*>(img);
  int nbias;
  if (kind == 0) {
    // [chunk c (4)][n (64)][8] with k = c*8+e = (i*3+j)*2 + ch   (i <-> dx+1, j <-> dy+1);
    // k = 18, 19 multiply a constant-1 input: the bias rounded to bf16 and the remainder of that rounding, so the GEMM
    // itself adds the bias to 2^-17 relative accuracy and the epilogue is a bare ReLU + pack
    for (int c = 0; c < 4; ++c)
      for (int n = 0; n < 64; ++n)
        for (int e = 0; e < 8; ++e) {
          const int k = c * 8 + e;
          float val = 0.0f;
          if (k < 18) { const int tap = k / 2, ch = k % 2; val = W(n, ch, tap / 3, tap % 3); }
          else if (k < 20 && n < cout) {
            const float bf = (float)bias[n];
            const uint32_t hb = (uint32_t)f32_to_bf16_rne(bf) << 16;
            float hi; memcpy(&hi, &hb, 4);
            val = (k == 18) ? hi : (bf - hi);
          }
          q[((size_t)c * 64 + n) * 8 + e] = f32_to_bf16_rne(val);
        }
    nbias = 64;
  } else {
    const int NB = (kind == 2) ? 16 : 64;
    // [dy (3)][chunk c8 (8)][n = blk*NB + co (3*NB)][8]; blk 0,1,2 <-> dx = +1,0,-1 <-> kernel row i = 2,1,0
    for (int dyi = 0; dyi < 3; ++dyi)
      for (int c8 = 0; c8 < 8; ++c8)
        for (int n = 0; n < 3 * NB; ++n)
          for (int e = 0; e < 8; ++e) {
            const int co = n % NB, blk = n / NB;
            q[(((size_t)dyi * 8 + c8) * (3 * NB) + n) * 8 + e] = f32_to_bf16_rne(W(co, c8 * 8 + e, 2 - blk, dyi));
          }
    nbias = NB;
  }
  const int64_t bias_off = (kind == 0) ? need - (int64_t)nbias * 4 : (kind == 1 ? conv_mid_extra_off() : conv_last_extra_off()) - (int64_t)nbias * 4;
  float* bq = reinterpret_cast<float*>(img + bias_off);
  for (int i = 0; i < nbias; ++i) bq[i] = (i < cout) ? (float)bias[i] : 0.0f;
  return need;
}

int mlsim_cnn_begin(mlsim_plan* pl, int32_t nlayers) {
  MLSIM_ENTER(pl);
  MLSIM_REQUIRE(nlayers >= 2 && nlayers <= 64, "the tensor-core path needs 2..64 conv layers");
  for (auto& l : pl->layers) if (l.d_wimg) { cudaFree(l.d_wimg); l.d_wimg = nullptr; }
  pl->layers.assign(nlayers, CnnLayer());
  pl->cnn_ready = false;
  drop_graphs(pl);                      // captured steps hold the old weight images
  return MLSIM_OK;
}

int mlsim_cnn_set_layer(mlsim_plan* pl, int32_t layer, int32_t cin, int32_t cout, int32_t ksize, const void* w,
                        const void* bias, int32_t src_dtype, int32_t relu_after) {
  if (!pl || !w || !bias) { set_error("null argument"); return MLSIM_EINVAL; }
  const int L = (int)pl->layers.size();
  MLSIM_REQUIRE(layer >= 0 && layer < L, "layer index out of range (call mlsim_cnn_begin first)");
  MLSIM_REQUIRE(ksize == 3, "only 3x3 / stride 1 / padding 1 convolutions are on the hot path");
  MLSIM_REQUIRE(src_dtype == 0 || src_dtype == 1, "src_dtype: 0 = f32, 1 = f64");
  if (layer == 0) MLSIM_REQUIRE(cin == 2 && cout >= 1 && cout <= 64, "first layer must be 2 -> (1..64) channels");
  else if (layer == L - 1) MLSIM_REQUIRE(cin >= 1 && cin <= 64 && cout >= 1 && cout <= 16, "last layer must be (1..64) -> (1..16) channels");
  else MLSIM_REQUIRE(cin >= 1 && cin <= 64 && cout >= 1 && cout <= 64, "hidden layers must be (1..64) -> (1..64) channels");
  if (layer != L - 1) MLSIM_REQUIRE(relu_after != 0, "ReLU after every hidden layer (reference CNN.forward / ResNetBlock.forward)");
  CnnLayer& l = pl->layers[layer];
  l.cin = cin; l.cout = cout; l.relu = relu_after ? 1 : 0;
  const size_t nw = (size_t)cout * cin * 9;
  l.w.resize(nw); l.b.resize(cout);
  if (src_dtype == 1) {
    memcpy(l.w.data(), w, nw * sizeof(double));
    memcpy(l.b.data(), bias, cout * sizeof(double));
  } else {
    for (size_t i = 0; i < nw; ++i) l.w[i] = ((const float*)w)[i];
    for (int i = 0; i < cout; ++i) l.b[i] = ((const float*)bias)[i];
  }
  l.set = true;
  l.res_mode = 0; l.res_src = -1; l.sw.clear(); l.sb.clear();
  pl->cnn_ready = false;
  drop_graphs(pl);                      // captured steps hold the old weight images
  return MLSIM_OK;
}

int mlsim_cnn_set_shortcut(mlsim_plan* pl, int32_t layer, int32_t src_layer, const void* w1x1, const void* b1x1,
                           int32_t src_dtype) {
  MLSIM_ENTER(pl);
  const int L = (int)pl->layers.size();
  MLSIM_REQUIRE(layer >= 1 && layer < L && pl->layers[layer].set, "set the layer before its shortcut (layer >= 1)");
  MLSIM_REQUIRE(src_layer >= -1 && src_layer < layer - 1, "the shortcut source must be an earlier layer's output (-1: network input)");
  MLSIM_REQUIRE(src_dtype == 0 || src_dtype == 1, "src_dtype: 0 = f32, 1 = f64");
  CnnLayer& l = pl->layers[layer];
  const int cs = (src_layer < 0) ? 2 : pl->layers[src_layer].cout;
  if (w1x1 == nullptr) {
    MLSIM_REQUIRE(src_layer >= 0 && layer < L - 1 && cs == l.cout, "identity shortcut: hidden layer, source with the same channel count");
    l.res_mode = 1;
  } else {
    MLSIM_REQUIRE(b1x1 != nullptr, "1x1 shortcut needs a bias");
    if (src_layer < 0) MLSIM_REQUIRE(layer < L - 1, "1x1 shortcut of the network input is supported on hidden layers");
    else MLSIM_REQUIRE(layer == L - 1 && l.cout <= 4, "1x1 shortcut of an activation is supported on the last layer (<= 4 output channels)");
    l.res_mode = (src_layer < 0) ? 2 : 3;
    const size_t nw = (size_t)l.cout * cs;
    l.sw.resize(nw); l.sb.resize(l.cout);
    for (size_t i = 0; i < nw; ++i) l.sw[i] = src_dtype == 1 ? ((const double*)w1x1)[i] : (double)((const float*)w1x1)[i];
    for (int i = 0; i < l.cout; ++i) l.sb[i] = src_dtype == 1 ? ((const double*)b1x1)[i] : (double)((const float*)b1x1)[i];
  }
  l.res_src = src_layer;
  pl->cnn_ready = false;
  drop_graphs(pl);                      // captured steps hold the old weight images
  return MLSIM_OK;
}

int mlsim_cnn_finalize(mlsim_plan* pl) {
  MLSIM_ENTER(pl);
  const int L = (int)pl->layers.size();
  MLSIM_REQUIRE(L >= 2, "no layers declared");
  for (int i = 0; i < L; ++i) {
    MLSIM_REQUIRE(pl->layers[i].set, "a layer was not set");
    if (i > 0) MLSIM_REQUIRE(pl->layers[i].cin == pl->layers[i - 1].cout, "channel mismatch between layers");
  }
  MLSIM_CUDA_CHECK(cudaSetDevice(pl->cfg.device));
  for (int li = 0; li < L; ++li) {
    CnnLayer& l = pl->layers[li];
    const int kind = (li == 0) ? 0 : (li == L - 1 ? 2 : 1);
    std::vector<uint8_t> img((size_t)mlsim_cnn_pack_host(kind, l.cin, l.cout, nullptr, nullptr, nullptr, 0));
    std::vector<double> bias = l.b;
    if (l.res_mode >= 2) for (int i = 0; i < l.cout; ++i) bias[i] += l.sb[i];     // the shortcut's bias rides on the layer's
    if (mlsim_cnn_pack_host(kind, l.cin, l.cout, l.w.data(), bias.data(), img.data(), (int64_t)img.size()) < 0) return MLSIM_EINVAL;
    if (l.res_mode >= 2) {
      // 1x1 weights, rounded to bf16 like every other conv weight (kept as fp32 in the image: CUDA-core epilogue math)
      float* ex = reinterpret_cast<float*>(img.data() + (kind == 1 ? conv_mid_extra_off() : conv_last_extra_off()));
      const int cs = (l.res_src < 0) ? 2 : pl->layers[l.res_src].cout;
      for (int co = 0; co < l.cout; ++co)
        for (int ci = 0; ci < cs; ++ci) {
          const uint16_t h = f32_to_bf16_rne((float)l.sw[(size_t)co * cs + ci]);
          const uint32_t bits = (uint32_t)h << 16;
          float f; memcpy(&f, &bits, 4);
          ex[(size_t)co * (kind == 1 ? 2 : 64) + ci] = f;
        }
    }
    if (l.d_wimg) { cudaFree(l.d_wimg); l.d_wimg = nullptr; }
    int rc = dev_alloc(pl, &l.d_wimg, img.size());
    if (rc) return rc;
    MLSIM_CUDA_CHECK(cudaMemcpy(l.d_wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  }
  // activation buffer assignment: the output of a layer must not overwrite its input nor a block input that a later
  // shortcut still needs (two buffers for a plain chain, three with shortcuts)
  {
    int need = 2;
    for (auto& l : pl->layers) if (l.res_mode == 1 || l.res_mode == 3) need = 3;
    std::vector<int> buf_of(L, -1);
    for (int li = 0; li < L; ++li) {
      CnnLayer& l = pl->layers[li];
      l.in_buf = li > 0 ? buf_of[li - 1] : -1;
      l.res_buf = (l.res_mode == 1 || l.res_mode == 3) ? buf_of[l.res_src] : -1;
      int pick = -1;
      for (int b = 0; b < need && pick < 0; ++b) {
        bool busy = (b == l.in_buf);
        for (int m = li; m < L && !busy; ++m) {           // block inputs still to be consumed (including by this layer)
          const CnnLayer& q = pl->layers[m];
          if ((q.res_mode == 1 || q.res_mode == 3) && q.res_src < li && buf_of[q.res_src] == b) busy = true;
        }
        if (!busy) pick = b;
      }
      MLSIM_REQUIRE(pick >= 0 || li == L - 1, "more than one outstanding shortcut source: not supported");
      l.out_buf = pick;
      buf_of[li] = pick;
    }
    if (need > pl->nact && pl->d_act[0]) {       // re-finalised with a deeper buffer requirement: reallocate
      for (int i = 0; i < 3; ++i) if (pl->d_act[i]) { cudaFree(pl->d_act[i]); pl->d_act[i] = nullptr; }
    }
    pl->nact = need;
  }
  if (!pl->d_act[0]) {
    const mlsim_config& c = pl->cfg;
    const int act_rows = pl->slab ? pl->nxl + 2 * MLSIM_SLAB_CNN_HALO : c.nx;
    pl->act_bytes = (size_t)c.batch * act_rows * 8 * (size_t)(c.ny + 2) * 16;
    for (int i = 0; i < pl->nact; ++i) {
      uint8_t* p = nullptr;
      int rc = dev_alloc(pl, &p, pl->act_bytes);
      if (rc) return rc;
      pl->d_act[i] = reinterpret_cast<__nv_bfloat16*>(p);
      MLSIM_CUDA_CHECK(cudaMemset(p, 0, pl->act_bytes));     // zero halo columns yp = 0 and yp = ny+1
    }
    pl->rs = conv_pick_strip(act_rows, c.ny, c.batch, pl->sm_count);
  }
  // fused first layer: layer 1 must be a hidden layer fed by layer 0 alone, and no shortcut may read layer 0's output
  {
    bool ok = pl->cfg.separate_first_layer == 0 && L >= 3;
    for (int li = 1; li < L && ok; ++li)
      if ((pl->layers[li].res_mode == 1 || pl->layers[li].res_mode == 3) && pl->layers[li].res_src == 0) ok = false;
    pl->fuse_first = ok;
  }
  MLSIM_CUDA_CHECK(cudaDeviceSynchronize());
  drop_graphs(pl);
  pl->cnn_ready = true;
  return MLSIM_OK;
}

int mlsim_step(mlsim_plan* pl, float* u, float* v, float* p_or_null, int32_t nsteps, int32_t apply_cnn,
               double cnn_input_scale, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v"))) return rc;
  if (p_or_null && (rc = check_field(p_or_null, "p"))) return rc;
  MLSIM_REQUIRE(nsteps >= 0, "nsteps must be >= 0");
  if (apply_cnn && !pl->cnn_ready) { set_error("apply_cnn set but the correction network is not finalised"); return MLSIM_ESTATE; }
  if (apply_cnn) MLSIM_REQUIRE(pl->layers.back().cout == 2, "v += model(v) needs a 2-channel network output (u, v)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rc2 = run_steps_graphed(pl, u, v, p_or_null, nsteps, apply_cnn, cnn_input_scale, st);
  if (!rc2) pl->fc_steps += nsteps;
  return rc2;
}

int mlsim_step_f64(mlsim_plan* pl, const double* u_in, const double* v_in, double* u_out, double* v_out, double* p_out,
                   int32_t nsteps, int32_t apply_cnn, double cnn_input_scale, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u_in, "u_in")) || (rc = check_field(v_in, "v_in")) || (rc = check_field(u_out, "u_out")) ||
      (rc = check_field(v_out, "v_out"))) return rc;
  if (p_out && (rc = check_field(p_out, "p_out"))) return rc;
  MLSIM_REQUIRE(nsteps >= 0, "nsteps must be >= 0");
  if (apply_cnn && !pl->cnn_ready) { set_error("apply_cnn set but the correction network is not finalised"); return MLSIM_ESTATE; }
  if (apply_cnn) MLSIM_REQUIRE(pl->layers.back().cout == 2, "v += model(v) needs a 2-channel network output (u, v)");
  const size_t n = pl->field_elems;
  if (!pl->d_fu) {
    if ((rc = dev_alloc(pl, &pl->d_fu, n)) || (rc = dev_alloc(pl, &pl->d_fv, n)) || (rc = dev_alloc(pl, &pl->d_fp, n))) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((rc = launch_cast_in(u_in, v_in, pl->d_fu, pl->d_fv, n, st))) return rc;
  if ((rc = run_steps_graphed(pl, pl->d_fu, pl->d_fv, p_out ? pl->d_fp : nullptr, nsteps, apply_cnn, cnn_input_scale, st)))
    return rc;
  pl->fc_steps += nsteps;
  return launch_cast_out(pl->d_fu, pl->d_fv, pl->d_fp, u_out, v_out, p_out, n, st);
}

int mlsim_rollout_host(mlsim_plan* pl, float* hu, float* hv, float* hp, int32_t nsteps, int32_t apply_cnn,
                       double cnn_input_scale) {
  if (!pl || !hu || !hv) { set_error("null argument"); return MLSIM_EINVAL; }
  MLSIM_PLAIN_ONLY(pl);
  MLSIM_REQUIRE(nsteps >= 0, "nsteps must be >= 0");
  if (apply_cnn && !pl->cnn_ready) { set_error("apply_cnn set but the correction network is not finalised"); return MLSIM_ESTATE; }
  if (apply_cnn) MLSIM_REQUIRE(pl->layers.back().cout == 2, "v += model(v) needs a 2-channel network output (u, v)");
  MLSIM_CUDA_CHECK(cudaSetDevice(pl->cfg.device));
  const size_t n = pl->field_elems;
  int rc;
  if (!pl->d_hu) {
    if ((rc = dev_alloc(pl, &pl->d_hu, n)) || (rc = dev_alloc(pl, &pl->d_hv, n)) || (rc = dev_alloc(pl, &pl->d_hp, n))) return rc;
    MLSIM_CUDA_CHECK(cudaStreamCreateWithFlags(&pl->in_stream, cudaStreamNonBlocking));
    MLSIM_CUDA_CHECK(cudaStreamCreateWithFlags(&pl->out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < mlsim_plan::MAX_CHUNKS; ++i) {
      MLSIM_CUDA_CHECK(cudaEventCreateWithFlags(&pl->ev_in[i], cudaEventDisableTiming));
      MLSIM_CUDA_CHECK(cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming));
    }
  }
  // Three-stage pipeline over blocks of independent trajectories: H2D (copy engine) | steps (SMs) | D2H (copy
  // engine).  Trajectories do not interact, so block k+1 uploads while block k computes and block k-1 downloads.
  const int B = pl->cfg.batch;
  // Block sizes: small first and last blocks (their upload / download is the only transfer time that is not hidden),
  // a large middle block (launch efficiency), all rounded to whole "rounds" of the persistent conv grid
  // (sm_count CTAs / work items per image) so that no block leaves most SMs idle in its last round.
  int bounds[mlsim_plan::MAX_CHUNKS + 1] = {0};
  int nchunks = B >= 16 ? 4 : (B >= 2 ? 2 : 1);
  for (int k = 0; k <= nchunks; ++k) bounds[k] = (int)((long)B * k / nchunks);
  if (B >= 24 && apply_cnn && pl->cnn_ready) {
    // one "round" = the images whose conv work items fill every persistent CTA exactly once
    const int items_per_img = ((pl->cfg.nx + pl->rs - 1) / pl->rs) * ((pl->cfg.ny + 127) / 128);
    const double round_imgs = (double)pl->sm_count / (items_per_img > 0 ? items_per_img : 1);
    const int total = (int)ceil(B / round_imgs - 1e-9);
    if (total >= 3) {
      // [head | the bulk | tail] in whole rounds.  The copy engines move a trajectory about twice as fast as the SMs
      // step it, so the bulk's upload hides behind a head of half its size, and the tail's download is the only
      // transfer left exposed at the end: for batches of 4-8 rounds [2 | bulk | 1] -- config 2: (18, 37, 9), measured
      // 1.99 ms per step against 2.01-2.07 for (9, 37, 18) and 2.19 for (9, 46, 9) (scripts/e2e_blocks.py); larger
      // batches keep [1 | bulk | a quarter].
      const bool small = total >= 4 && total <= 8;
      int last = small ? 1 : (total + 2) / 4;
      if (last < 1) last = 1;
      const int first = small ? 2 : 1;
      const int e1 = (int)floor(first * round_imgs + 1e-9), e2 = (int)floor((total - last) * round_imgs + 1e-9);
      if (e1 >= 1 && e2 > e1 && e2 < B) { nchunks = 3; bounds[0] = 0; bounds[1] = e1; bounds[2] = e2; bounds[3] = B; }
    }
  }
  if (pl->host_nsizes > 0) {                                   // caller's choice: explicit block sizes
    nchunks = pl->host_nsizes;
    for (int k = 0; k < nchunks; ++k) bounds[k + 1] = bounds[k] + pl->host_sizes[k];
  } else if (pl->cfg.host_blocks > 0) {                        // caller's choice: equal blocks
    const int v = pl->cfg.host_blocks < B ? pl->cfg.host_blocks : B;
    nchunks = v;
    for (int k = 0; k <= v; ++k) bounds[k] = (int)((long)B * k / v);
  }
  const size_t img = (size_t)pl->cfg.nx * pl->cfg.ny;
  cudaStream_t sc = pl->own_stream;
  for (int k = 0; k < nchunks; ++k) {
    const int b0 = bounds[k], b1 = bounds[k + 1];
    const size_t off = (size_t)b0 * img, cnt = (size_t)(b1 - b0) * img;
    MLSIM_CUDA_CHECK(cudaMemcpyAsync(pl->d_hu + off, hu + off, cnt * sizeof(float), cudaMemcpyHostToDevice, pl->in_stream));
    MLSIM_CUDA_CHECK(cudaMemcpyAsync(pl->d_hv + off, hv + off, cnt * sizeof(float), cudaMemcpyHostToDevice, pl->in_stream));
    MLSIM_CUDA_CHECK(cudaEventRecord(pl->ev_in[k], pl->in_stream));
    MLSIM_CUDA_CHECK(cudaStreamWaitEvent(sc, pl->ev_in[k], 0));
    if ((rc = run_steps(pl, pl->d_hu + off, pl->d_hv + off, hp ? pl->d_hp + off : nullptr, nsteps, apply_cnn,
                        cnn_input_scale, sc, Chunk{b0, b1 - b0}))) return rc;
    MLSIM_CUDA_CHECK(cudaEventRecord(pl->ev_done[k], sc));
    MLSIM_CUDA_CHECK(cudaStreamWaitEvent(pl->out_stream, pl->ev_done[k], 0));
    MLSIM_CUDA_CHECK(cudaMemcpyAsync(hu + off, pl->d_hu + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, pl->out_stream));
    MLSIM_CUDA_CHECK(cudaMemcpyAsync(hv + off, pl->d_hv + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, pl->out_stream));
    if (hp) MLSIM_CUDA_CHECK(cudaMemcpyAsync(hp + off, pl->d_hp + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, pl->out_stream));
  }
  MLSIM_CUDA_CHECK(cudaStreamSynchronize(pl->out_stream));
  MLSIM_CUDA_CHECK(cudaStreamSynchronize(sc));
  pl->fc_steps += nsteps;
  return MLSIM_OK;
}

int mlsim_set_host_blocks(mlsim_plan* pl, const int32_t* sizes, int32_t n) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  if (sizes == nullptr || n == 0) { pl->host_nsizes = 0; return MLSIM_OK; }
  MLSIM_REQUIRE(n >= 1 && n <= mlsim_plan::MAX_CHUNKS, "at most 8 blocks");
  int sum = 0;
  for (int k = 0; k < n; ++k) { MLSIM_REQUIRE(sizes[k] >= 1, "block sizes must be >= 1"); sum += sizes[k]; }
  MLSIM_REQUIRE(sum == pl->cfg.batch, "block sizes must add up to the batch");
  for (int k = 0; k < n; ++k) pl->host_sizes[k] = sizes[k];
  pl->host_nsizes = n;
  return MLSIM_OK;
}

int mlsim_explicit_terms(mlsim_plan* pl, const float* u, const float* v, float* du, float* dv, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v")) || (rc = check_field(du, "du")) || (rc = check_field(dv, "dv"))) return rc;
  StageArgs a;
  a.us = u; a.vs = v; a.u0 = u; a.v0 = v; a.accu = nullptr; a.accv = nullptr; a.outu = du; a.outv = dv;
  a.ca = 0.f; a.cs = 0.f; a.mode = 0;
  return launch_explicit(a, pl->sp, pl->d_forcing, reinterpret_cast<cudaStream_t>(stream));
}

int mlsim_project(mlsim_plan* pl, float* u, float* v, float* p_or_null, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v"))) return rc;
  return project(pl, u, v, p_or_null, reinterpret_cast<cudaStream_t>(stream), Chunk{0, pl->cfg.batch});
}

// out = irfft2(rfft2(in) * table): the spectral filter of the initial-condition generator (torch_cfd's
// filtered_velocity_field, reference src/inference.py:76-78) on the projection's own transform kernels.
// table: HOST fp64 [nx][ny/2+1] in natural kx order.  Set-up path: synchronises `stream` (table upload).
int mlsim_spectral_multiply(mlsim_plan* pl, const double* table, const float* in, float* out, void* stream) {
  if (!pl || !table) { set_error("null argument"); return MLSIM_EINVAL; }
  MLSIM_CUDA_CHECK(cudaSetDevice(pl->cfg.device));
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(in, "in")) || (rc = check_field(out, "out"))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nx = pl->cfg.nx, nyc = pl->sp.nyc, r0 = pl->r0, n2 = pl->n2len;
  std::vector<float> t((size_t)nx * nyc);
  const double norm = 1.0 / ((double)nx * pl->cfg.ny);
  for (int k1 = 0; k1 < r0; ++k1)
    for (int k2 = 0; k2 < n2; ++k2) {
      const int kx = k1 + r0 * k2;
      for (int k = 0; k < nyc; ++k) t[((size_t)k1 * n2 + k2) * nyc + k] = (float)(table[(size_t)kx * nyc + k] * norm);
    }
  if (!pl->d_filt && (rc = dev_alloc(pl, &pl->d_filt, t.size()))) return rc;
  MLSIM_CUDA_CHECK(cudaStreamSynchronize(st));            // an earlier call may still be reading the table
  MLSIM_CUDA_CHECK(cudaMemcpy(pl->d_filt, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
  SolverParams sp = pl->sp;
  sp.raw = 1;
  if ((rc = launch_div_rfft(in, in, pl->d_spec, pl->d_twy, sp, st))) return rc;
  if (pl->r0 == 1) rc = launch_col_solve(pl->d_spec, pl->d_filt, pl->d_twx, sp, sp.nx, sp.batch, 1, st);
  else rc = launch_col_solve_split(pl->d_spec, pl->d_split, pl->d_filt, pl->d_twx, pl->d_tw_n2, sp, pl->r0, st);
  if (rc) return rc;
  return launch_irfft_grad(pl->d_spec, in, in, out, out, out, pl->d_twy, sp, st);
}

// Per trajectory: u, v *= max_velocity / max sqrt(u^2 + v^2) -- the rescaling step of filtered_velocity_field.
int mlsim_normalize_speed(mlsim_plan* pl, float* u, float* v, double max_velocity, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v"))) return rc;
  MLSIM_REQUIRE(max_velocity > 0, "max_velocity must be positive");
  if (!pl->d_smax && (rc = dev_alloc(pl, &pl->d_smax, (size_t)pl->cfg.batch))) return rc;
  return launch_normalize_speed(u, v, pl->d_smax, pl->cfg.batch, pl->cfg.nx * pl->cfg.ny, (float)max_velocity,
                                reinterpret_cast<cudaStream_t>(stream));
}

int mlsim_divergence(mlsim_plan* pl, const float* u, const float* v, float* d, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v")) || (rc = check_field(d, "div"))) return rc;
  return launch_divergence(u, v, d, pl->sp, reinterpret_cast<cudaStream_t>(stream));
}

int mlsim_cnn_forward(mlsim_plan* pl, const float* u, const float* v, float* y, double scale, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v")) || (rc = check_field(y, "y"))) return rc;
  return cnn_apply(pl, u, v, nullptr, nullptr, y, scale, reinterpret_cast<cudaStream_t>(stream), Chunk{0, pl->cfg.batch});
}

int mlsim_cnn_correct(mlsim_plan* pl, float* u, float* v, double scale, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v"))) return rc;
  if (!pl->cnn_ready) { set_error("correction network not finalised"); return MLSIM_ESTATE; }
  MLSIM_REQUIRE(pl->layers.back().cout == 2, "the fused correction add needs a 2-channel output (u, v)");
  return cnn_apply(pl, u, v, u, v, nullptr, scale, reinterpret_cast<cudaStream_t>(stream), Chunk{0, pl->cfg.batch});
}

// ---------------------------------------------------------------------------------------------------------------
// Slab decomposition along x (BASELINE config 5: one large grid over several GPUs).  The caller (one process per GPU)
// owns every buffer that is exchanged and runs the exchanges itself (NCCL through torch.distributed, see slab.py):
//   field pair  uv[2][batch][ext_rows][ny] fp32, ext_rows = nxl + 2*8: local row i lives at ext row i + 8
//   forward  spectral chunks [world][batch][nxl  ][kywp] float2   (all-to-all: chunk r -> rank r)
//   backward spectral chunks [world][batch][nxl+1][kywp] float2   (row nxl = x-halo of p-hat)
// One RK stage =  halo exchange(us: 3 rows up, 2 rows down) -> stage_a -> all-to-all -> stage_b -> all-to-all -> stage_c.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct UV { float* u; float* v; };
UV uv_origin(const mlsim_plan* pl, float* base) {
  const size_t row0 = (size_t)MLSIM_SLAB_HALO * pl->cfg.ny;
  return UV{base + row0, base + (size_t)pl->cfg.batch * pl->ext_rows * pl->cfg.ny + row0};
}
}  // namespace

int mlsim_slab_geometry(mlsim_plan* pl, int64_t* out8) {
  if (!pl || !out8) { set_error("null argument"); return MLSIM_EINVAL; }
  MLSIM_SLAB_ONLY(pl);
  out8[0] = pl->nxl; out8[1] = MLSIM_SLAB_HALO; out8[2] = pl->ext_rows; out8[3] = pl->kyw; out8[4] = pl->kywp;
  out8[5] = pl->kyv; out8[6] = pl->r0; out8[7] = MLSIM_SLAB_CNN_HALO;
  return MLSIM_OK;
}

int mlsim_slab_stage_a(mlsim_plan* pl, int32_t stage, const float* us, const float* u0, float* acc, float* out,
                       void* spec_send, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  MLSIM_REQUIRE(stage >= 0 && stage <= 3, "stage must be 0..3");
  int rc;
  if ((rc = check_field(us, "us")) || (rc = check_field(u0, "u0")) || (rc = check_field(acc, "acc")) ||
      (rc = check_field(out, "out"))) return rc;
  if (spec_send == nullptr) MLSIM_REQUIRE(pl->have_peers, "spec_send is NULL and no peer buffers were registered (mlsim_slab_set_peers)");
  else if ((rc = check_field(spec_send, "spec_send"))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float dt = (float)pl->cfg.dt;
  const UV s = uv_origin(pl, const_cast<float*>(us)), b = uv_origin(pl, const_cast<float*>(u0)), a = uv_origin(pl, acc),
           o = uv_origin(pl, out);
  StageArgs g;
  g.us = s.u; g.vs = s.v; g.u0 = b.u; g.v0 = b.v; g.accu = a.u; g.accv = a.v; g.outu = o.u; g.outv = o.v;
  switch (stage) {                                   // coefficients as in solver_step()
    case 0: g.ca = dt / 6.0f; g.cs = 0.5f * dt; g.mode = 1; break;
    case 1: g.ca = dt / 3.0f; g.cs = 0.5f * dt; g.mode = 2; break;
    case 2: g.ca = dt / 3.0f; g.cs = dt; g.mode = 2; break;
    default: g.ca = 0.0f; g.cs = dt / 6.0f; g.mode = 3; break;
  }
  { ProfScope ps(pl, 0, st); if ((rc = launch_explicit(g, pl->sp, pl->d_forcing, st))) return rc; }
  ProfScope ps(pl, 1, st);
  return launch_div_rfft(o.u, o.v, static_cast<float2*>(spec_send), pl->d_twy, pl->sp, st,
                         spec_send ? nullptr : &pl->fwd_peers, pl->cfg.slab_world);
}

int mlsim_slab_divergence_rfft(mlsim_plan* pl, const float* uv, void* spec_send, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  int rc;
  if ((rc = check_field(uv, "uv"))) return rc;
  if (spec_send == nullptr) MLSIM_REQUIRE(pl->have_peers, "spec_send is NULL and no peer buffers were registered (mlsim_slab_set_peers)");
  else if ((rc = check_field(spec_send, "spec_send"))) return rc;
  const UV o = uv_origin(pl, const_cast<float*>(uv));
  return launch_div_rfft(o.u, o.v, static_cast<float2*>(spec_send), pl->d_twy, pl->sp, reinterpret_cast<cudaStream_t>(stream),
                         spec_send ? nullptr : &pl->fwd_peers, pl->cfg.slab_world);
}

int mlsim_slab_stage_b(mlsim_plan* pl, const void* spec_recv, void* spec_send2, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  int rc;
  if ((rc = check_field(spec_recv, "spec_recv"))) return rc;
  if (spec_send2 == nullptr) MLSIM_REQUIRE(pl->have_peers, "spec_send2 is NULL and no peer buffers were registered (mlsim_slab_set_peers)");
  else if ((rc = check_field(spec_send2, "spec_send2"))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope ps(pl, 2, st);
  return launch_col_solve_slab(static_cast<const float2*>(spec_recv), static_cast<float2*>(spec_send2), pl->d_split,
                               pl->d_inv_eig, pl->d_twx, pl->d_tw_n2, pl->sp, pl->nxg, pl->cfg.slab_world, pl->kyv,
                               pl->r0, st, spec_send2 ? nullptr : &pl->bwd_peers);
}

int mlsim_slab_stage_c(mlsim_plan* pl, const void* spec_recv2, float* uv, float* p_or_null, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  int rc;
  if ((rc = check_field(spec_recv2, "spec_recv2")) || (rc = check_field(uv, "uv"))) return rc;
  if (p_or_null && (rc = check_field(p_or_null, "p"))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const UV o = uv_origin(pl, uv);
  ProfScope ps(pl, 3, st);
  return launch_irfft_grad(static_cast<const float2*>(spec_recv2), o.u, o.v, o.u, o.v, p_or_null, pl->d_twy, pl->sp, st,
                           pl->cfg.slab_world);
}

// Peer-memory halo exchange: copy this rank's top `lo` rows into the upper neighbour's low halo and its bottom `hi`
// rows into the lower neighbour's high halo.  up_uv / dn_uv: the neighbours' field pair (same buffer role) as mapped
// into this process, or NULL to skip that side (edges of a non-periodic exchange).  The caller follows with a barrier.
int mlsim_slab_halo_push(mlsim_plan* pl, const float* uv, float* up_uv, float* dn_uv, int32_t lo, int32_t hi, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  int rc;
  if ((rc = check_field(uv, "uv"))) return rc;
  if (up_uv && (rc = check_field(up_uv, "up_uv"))) return rc;
  if (dn_uv && (rc = check_field(dn_uv, "dn_uv"))) return rc;
  MLSIM_REQUIRE(lo >= 0 && hi >= 0 && lo <= MLSIM_SLAB_HALO && hi <= MLSIM_SLAB_HALO && lo <= pl->nxl && hi <= pl->nxl, "halo width out of range");
  if ((!up_uv || lo == 0) && (!dn_uv || hi == 0)) return MLSIM_OK;
  return launch_halo_push(uv, lo > 0 ? up_uv : nullptr, hi > 0 ? dn_uv : nullptr, 2 * pl->cfg.batch, pl->ext_rows, pl->cfg.ny,
                          pl->nxl, MLSIM_SLAB_HALO, lo, hi, reinterpret_cast<cudaStream_t>(stream));
}

// Peer-memory spectral exchange: fwd_recv[r] / bwd_recv[r] = base of rank r's forward / backward receive buffer as
// mapped into THIS process (CUDA IPC / symmetric memory over NVLink).  Afterwards stage_a (spec_send == NULL) stores
// column chunk r straight into rank r's buffer at this rank's slot, and stage_b (spec_send2 == NULL) does the same for
// the backward chunks; the caller replaces each all-to-all by a cross-rank barrier.
int mlsim_slab_set_peers(mlsim_plan* pl, const uint64_t* fwd_recv, const uint64_t* bwd_recv) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  if (!fwd_recv || !bwd_recv) { pl->have_peers = false; return MLSIM_OK; }
  const int W = pl->cfg.slab_world, r = pl->cfg.slab_rank;
  const size_t cf = (size_t)pl->cfg.batch * pl->nxl * pl->kywp, cb = (size_t)pl->cfg.batch * (pl->nxl + 1) * pl->kywp;
  for (int i = 0; i < MLSIM_MAX_SLAB_RANKS; ++i) {
    const int j = i < W ? i : 0;
    MLSIM_REQUIRE(fwd_recv[j] != 0 && bwd_recv[j] != 0 && (fwd_recv[j] & 15u) == 0 && (bwd_recv[j] & 15u) == 0,
                  "peer buffer pointers must be non-null and 16-byte aligned");
    pl->fwd_peers.p[i] = reinterpret_cast<float2*>(fwd_recv[j]) + (size_t)r * cf;
    pl->bwd_peers.p[i] = reinterpret_cast<float2*>(bwd_recv[j]) + (size_t)r * cb;
  }
  pl->have_peers = true;
  return MLSIM_OK;
}

// v += model(v) on the slab.  The CNN pads with ZEROS at the edges of the global grid (Conv2d padding=1, reference
// src/models/tools.py), so it is not periodic: the 7 conv layers need 7 rows of state from each interior neighbour
// (halo rows [-7, 0) and [nxl, nxl+7) must be current), the network runs on rows [lo, hi) = the slab plus those
// halos (the outermost l rows of layer l's output are contaminated by the artificial edge and never reach local rows).
int mlsim_slab_cnn_correct(mlsim_plan* pl, float* uv, double scale, void* stream) {
  MLSIM_ENTER(pl);
  MLSIM_SLAB_ONLY(pl);
  int rc;
  if ((rc = check_field(uv, "uv"))) return rc;
  if (!pl->cnn_ready) { set_error("correction network not finalised"); return MLSIM_ESTATE; }
  MLSIM_REQUIRE(pl->layers.back().cout == 2, "the fused correction add needs a 2-channel output (u, v)");
  MLSIM_REQUIRE((int)pl->layers.size() <= MLSIM_SLAB_CNN_HALO, "slab mode supports at most 7 conv layers (halo depth)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const mlsim_config& c = pl->cfg;
  const int W = c.slab_world, r = c.slab_rank;
  const int lo = (r == 0) ? 0 : -MLSIM_SLAB_CNN_HALO;
  const int hi = pl->nxl + ((r == W - 1) ? 0 : MLSIM_SLAB_CNN_HALO);
  const int rows = hi - lo;
  const UV o = uv_origin(pl, uv);
  float* u = o.u + (ptrdiff_t)lo * c.ny;
  float* v = o.v + (ptrdiff_t)lo * c.ny;
  return run_cnn_layers(pl, u, v, u, v, nullptr, scale, rows, c.batch, pl->sp.img_stride, 0, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Data generation (reference src/data_generation.py:193-213): one stored pair =
//   n_between fine steps; coarse_t = downsample(fine); n_inner fine steps; one coarse step of coarse_t
// all stream-ordered on the device, nothing returns to the host between the pieces.
// ---------------------------------------------------------------------------------------------------------------
int mlsim_downsample_staggered(const float* u, const float* v, float* cu, float* cv, int32_t batch, int32_t nx,
                               int32_t ny, int32_t factor, void* stream) {
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v")) || (rc = check_field(cu, "cu")) || (rc = check_field(cv, "cv"))) return rc;
  MLSIM_REQUIRE(batch >= 1 && nx >= 1 && ny >= 1 && factor >= 1, "bad shape");
  {
    cudaPointerAttributes at;                      // no plan here: launch on the device that owns the fields
    MLSIM_CUDA_CHECK(cudaPointerGetAttributes(&at, u));
    MLSIM_REQUIRE(at.type == cudaMemoryTypeDevice, "u must be device memory");
    MLSIM_CUDA_CHECK(cudaSetDevice(at.device));
  }
  MLSIM_REQUIRE(nx % factor == 0 && ny % factor == 0, "`factor` must divide the grid (block_reduce of the reference raises)");
  MLSIM_REQUIRE(factor == 1 || ((ny % 4) == 0), "ny must be a multiple of 4");
  return launch_downsample_staggered(u, v, cu, cv, batch, nx, ny, factor, reinterpret_cast<cudaStream_t>(stream));
}

int mlsim_datagen_pair(mlsim_plan* fine, mlsim_plan* coarse, float* u, float* v, float* cu, float* cv,
                       int32_t n_between, int32_t n_inner, void* stream) {
  if (!fine || !coarse) { set_error("null plan"); return MLSIM_EINVAL; }
  MLSIM_CUDA_CHECK(cudaSetDevice(fine->cfg.device));
  MLSIM_PLAIN_ONLY(fine);
  MLSIM_PLAIN_ONLY(coarse);
  int rc;
  if ((rc = check_field(u, "u")) || (rc = check_field(v, "v")) || (rc = check_field(cu, "cu")) || (rc = check_field(cv, "cv"))) return rc;
  const mlsim_config& f = fine->cfg;
  const mlsim_config& c = coarse->cfg;
  MLSIM_REQUIRE(n_between >= 0 && n_inner >= 0, "step counts must be >= 0");
  MLSIM_REQUIRE(f.batch == c.batch && f.device == c.device, "fine and coarse plans must share batch and device");
  MLSIM_REQUIRE(f.nx % c.nx == 0 && f.ny % c.ny == 0 && f.nx / c.nx == f.ny / c.ny, "coarse grid must divide the fine grid isotropically");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Chunk fc{0, f.batch}, cc{0, c.batch};
  if ((rc = run_steps(fine, u, v, nullptr, n_between, 0, 1.0, st, fc))) return rc;
  if ((rc = launch_downsample_staggered(u, v, cu, cv, f.batch, f.nx, f.ny, f.nx / c.nx, st))) return rc;
  if ((rc = run_steps(fine, u, v, nullptr, n_inner, 0, 1.0, st, fc))) return rc;
  return run_steps(coarse, cu, cv, nullptr, 1, 0, 1.0, st, cc);
}

int mlsim_finite_check_enable(mlsim_plan* pl, int32_t every) {
  MLSIM_ENTER(pl);
  MLSIM_PLAIN_ONLY(pl);
  MLSIM_REQUIRE(every >= 0, "every must be >= 0");
  if (every > 0 && !pl->d_fcflag) {
    int rc = dev_alloc(pl, &pl->d_fcflag, 1);
    if (rc) return rc;
    MLSIM_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&pl->h_fcflag), sizeof(unsigned long long), cudaHostAllocDefault));
  }
  if (pl->d_fcflag) {
    MLSIM_CUDA_CHECK(cudaDeviceSynchronize());                 // set-up call: no check of an earlier run is in flight
    MLSIM_CUDA_CHECK(cudaMemset(pl->d_fcflag, 0xff, sizeof(unsigned long long)));
    *pl->h_fcflag = ~0ull;
  }
  pl->fc_every = every;
  pl->fc_steps = 0;
  return MLSIM_OK;
}

int mlsim_finite_check_read(mlsim_plan* pl, int64_t* first_bad_step, int64_t* steps_issued) {
  if (!pl || !first_bad_step || !steps_issued) { set_error("null argument"); return MLSIM_EINVAL; }
  const unsigned long long f = pl->h_fcflag ? *reinterpret_cast<volatile unsigned long long*>(pl->h_fcflag) : ~0ull;
  *first_bad_step = (f == ~0ull) ? -1 : (int64_t)f;
  *steps_issued = pl->fc_steps;
  return MLSIM_OK;
}

int64_t mlsim_step_graph_launches(mlsim_plan* pl) { return pl ? pl->graph_launches : -1; }

int mlsim_launches_per_step(mlsim_plan* pl, int32_t apply_cnn) {
  if (!pl) return MLSIM_EINVAL;
  int n = pl->cluster_projection ? 8 : 16;   // 4 stages x (explicit + cluster projection | explicit, div+rfft, column solve, irfft+grad)
  if (pl->slab || pl->r0 > 1) n = 4 * 6;     // split x-transform: explicit, div+rfft, pass A, column solve, pass C, irfft+grad
  if (!pl->slab) n *= solver_blocks(pl, apply_cnn && pl->cnn_ready, pl->cfg.batch);   // one set of solver launches per batch block
  if (apply_cnn && pl->cnn_ready) n += (int)pl->layers.size() - (pl->fuse_first ? 1 : 0);
  return n;
}

int64_t mlsim_workspace_bytes(mlsim_plan* pl) { return pl ? pl->ws_bytes : 0; }

static int conv_debug_read(mlsim_plan* pl, double* out8, int first_cta) {
  if (!pl || !out8 || !pl->d_dbg) { set_error("no conv debug counters (run mlsim_time_kernel(5|6|8) first)"); return MLSIM_EINVAL; }
  std::vector<unsigned long long> h((size_t)1024 * 8);
  MLSIM_CUDA_CHECK(cudaMemcpy(h.data(), pl->d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  int n = 0;
  for (int k = 0; k < 8; ++k) out8[k] = 0.0;
  for (int b = first_cta; b < first_cta + 512; ++b) {
    if (h[(size_t)b * 8] == 0) continue;
    ++n;
    for (int k = 0; k < 8; ++k) out8[k] += (double)h[(size_t)b * 8 + k];
  }
  for (int k = 0; k < 8; ++k) out8[k] /= (n ? n : 1);
  return n;
}
int mlsim_conv_debug(mlsim_plan* pl, double* out8) { return conv_debug_read(pl, out8, 0); }
int mlsim_conv_debug_first(mlsim_plan* pl, double* out8) { return conv_debug_read(pl, out8, 512); }

int mlsim_profile_enable(mlsim_plan* pl, int32_t kernel_class, int32_t max_samples) {
  MLSIM_ENTER(pl);
  MLSIM_REQUIRE(kernel_class >= -1 && kernel_class <= 8 && max_samples >= 0 && max_samples <= (1 << 20), "bad arguments");
  MLSIM_CUDA_CHECK(cudaSetDevice(pl->cfg.device));
  while ((int)pl->prof_events.size() < max_samples) {
    cudaEvent_t a, b;
    MLSIM_CUDA_CHECK(cudaEventCreate(&a));
    MLSIM_CUDA_CHECK(cudaEventCreate(&b));
    pl->prof_events.emplace_back(a, b);
  }
  pl->prof_class = kernel_class;
  pl->prof_used = 0;
  return MLSIM_OK;
}

int mlsim_profile_read(mlsim_plan* pl, float* total_ms, int32_t* count) {
  if (!pl || !total_ms || !count) { set_error("null argument"); return MLSIM_EINVAL; }
  MLSIM_CUDA_CHECK(cudaSetDevice(pl->cfg.device));
  float tot = 0.f;
  for (size_t i = 0; i < pl->prof_used; ++i) {
    MLSIM_CUDA_CHECK(cudaEventSynchronize(pl->prof_events[i].second));
    float ms = 0.f;
    MLSIM_CUDA_CHECK(cudaEventElapsedTime(&ms, pl->prof_events[i].first, pl->prof_events[i].second));
    tot += ms;
  }
  *total_ms = tot;
  *count = (int32_t)pl->prof_used;
  pl->prof_used = 0;
  return MLSIM_OK;
}

int mlsim_time_kernel(mlsim_plan* pl, int32_t which, int32_t warmup, int32_t iters, float* ms_out, void* stream) {
  if (!pl || !ms_out) { set_error("null argument"); return MLSIM_EINVAL; }
  MLSIM_CUDA_CHECK(cudaSetDevice(pl->cfg.device));
  MLSIM_PLAIN_ONLY(pl);
  MLSIM_REQUIRE(which >= 0 && which <= 8 && iters >= 1 && warmup >= 0, "bad arguments");
  if (which == 7 && !pl->cluster_projection) { set_error("cluster projection not available for this grid"); return MLSIM_EINVAL; }
  if (which >= 4 && !pl->cnn_ready) { set_error("correction network not finalised"); return MLSIM_ESTATE; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = pl->field_elems;
  int rc;
  if (!pl->d_tu) {
    if ((rc = dev_alloc(pl, &pl->d_tu, n)) || (rc = dev_alloc(pl, &pl->d_tv, n)) || (rc = dev_alloc(pl, &pl->d_dbg, (size_t)1024 * 8))) return rc;
    MLSIM_CUDA_CHECK(cudaMemsetAsync(pl->d_dbg, 0, 1024 * 8 * sizeof(unsigned long long), st));
    MLSIM_CUDA_CHECK(cudaMemsetAsync(pl->d_tu, 0, n * sizeof(float), st));
    MLSIM_CUDA_CHECK(cudaMemsetAsync(pl->d_tv, 0, n * sizeof(float), st));
  }
  const mlsim_config& c = pl->cfg;
  const float dt = (float)c.dt;
  auto run = [&]() -> int {
    switch (which) {
      case 0: {
        StageArgs a;
        a.us = pl->d_tu; a.vs = pl->d_tv; a.u0 = pl->d_w2u; a.v0 = pl->d_w2v; a.accu = pl->d_accu; a.accv = pl->d_accv;
        a.outu = pl->d_w1u; a.outv = pl->d_w1v; a.ca = 0.f * dt; a.cs = 0.f * dt; a.mode = 2;
        return launch_explicit(a, pl->sp, pl->d_forcing, st);
      }
      case 7: return launch_project_cluster(pl->d_w1u, pl->d_w1v, nullptr, pl->d_inv_eig, pl->d_twy, pl->sp, st);
      case 1: return launch_div_rfft(pl->d_tu, pl->d_tv, pl->d_spec, pl->d_twy, pl->sp, st);
      case 2:
        if (pl->r0 == 1) return launch_col_solve(pl->d_spec, pl->d_inv_eig, pl->d_twx, pl->sp, pl->sp.nx, pl->sp.batch, 1, st);
        return launch_col_solve_split(pl->d_spec, pl->d_split, pl->d_inv_eig, pl->d_twx, pl->d_tw_n2, pl->sp, pl->r0, st);
      case 3: return launch_irfft_grad(pl->d_spec, pl->d_tu, pl->d_tv, pl->d_w1u, pl->d_w1v, nullptr, pl->d_twy, pl->sp, st);
      case 4: {
        ConvFirstArgs f;
        f.u = pl->d_tu; f.v = pl->d_tv; f.out = pl->d_act[0]; f.wimg = pl->layers[0].d_wimg; f.scale = 1.0f;
        f.nx = c.nx; f.ny = c.ny; f.batch = c.batch; f.reverse = 0; f.img_stride = (long long)c.nx * c.ny;
        return launch_conv_first(f, pl->sm_count, st);
      }
      default: {
        const int L = (int)pl->layers.size();
        const int l = (which == 5 || which == 8) ? 1 : L - 1;
        if ((which == 5 || which == 8) && L < 3) { set_error("no mid layer in this network"); return MLSIM_EINVAL; }
        ConvMidArgs m;
        m.in = pl->d_act[0]; m.out = pl->d_act[1]; m.u = pl->d_w1u; m.v = pl->d_w1v; m.y_nchw = nullptr;
        m.wimg = pl->layers[l].d_wimg; m.nx = c.nx; m.ny = c.ny; m.batch = c.batch; m.rs = pl->rs;
        m.img_stride = (long long)c.nx * c.ny; m.res_mode = 0; m.res = nullptr; m.sscale = 1.0f;
        m.cout = pl->layers[l].cout; m.relu = pl->layers[l].relu; m.dbg = pl->d_dbg; m.reverse = 0;
        m.fu = pl->d_tu; m.fv = pl->d_tv; m.wimg0 = pl->layers[0].d_wimg; m.fscale = 1.0f;
        if (which == 8) return launch_conv_mid_fused(m, pl->sm_count, st);
        return (which == 5) ? launch_conv_mid(m, pl->sm_count, st) : launch_conv_last(m, pl->sm_count, st);
      }
    }
  };
  for (int i = 0; i < warmup; ++i) if ((rc = run())) return rc;
  cudaEvent_t e0, e1;
  MLSIM_CUDA_CHECK(cudaEventCreate(&e0));
  MLSIM_CUDA_CHECK(cudaEventCreate(&e1));
  MLSIM_CUDA_CHECK(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) if ((rc = run())) return rc;
  MLSIM_CUDA_CHECK(cudaEventRecord(e1, st));
  MLSIM_CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  MLSIM_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  return MLSIM_OK;
}

}  // extern "C"

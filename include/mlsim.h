/*
 * mlsim.h -- C ABI of the B200-native rollout hot path (solver step + learned correction).
 *
 * The reference (BigBalloon8/ml-accelerated-simulation) is pure Python and has no FFI; the boundary it
 * exposes for this path is the Python call surface of torch_cfd plus src/models.  Each entry point
 * below names the reference call it stands in for.  Conventions:
 *   - return 0 on success, a negative MLSIM_E* code otherwise; no exceptions cross the ABI;
 *     mlsim_last_error() returns a thread-local message for the last failure;
 *   - field pointers are DEVICE pointers to contiguous fp32 arrays [batch][nx][ny] (axis 0 = x,
 *     axis 1 = y contiguous: the reference's GridVariable.data layout, src/data_generation.py:219-220),
 *     16-byte aligned, owned by the caller; everything else is owned by the plan;
 *   - all device work is enqueued on the cudaStream_t passed as `stream` (void*), stream-ordered, with
 *     no hidden synchronisation, except the *_host entry points which synchronise before returning;
 *   - a plan is not thread-safe; distinct plans are independent.  One plan == one device.
 */
#ifndef MLSIM_H
#define MLSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLSIM_OK            0
#define MLSIM_EINVAL       -1   /* bad argument / unsupported configuration */
#define MLSIM_ECUDA        -2   /* CUDA runtime error (message has the detail) */
#define MLSIM_ENOMEM       -3
#define MLSIM_ESTATE       -4   /* call out of order (e.g. cnn not finalised) */

#define MLSIM_ADV_VAN_LEER      0
#define MLSIM_ADV_UPWIND        1
#define MLSIM_ADV_LAX_WENDROFF  2

typedef struct mlsim_plan mlsim_plan;

/* Geometry + physics of one rollout.  Stands in for the objects built at src/inference.py:54-93:
 * grids.Grid((nx,ny), domain=((0,lx),(0,ly))), stable_time_step(...) -> dt,
 * NavierStokes2DFVMProjection(viscosity, density, drag, forcing=KolmogorovForcing(wave_number, scale)). */
typedef struct {
  int32_t nx, ny, batch;          /* powers of two, 64 <= n <= 8192 (batch >= 1); nx is the GLOBAL extent in slab mode */
  double  lx, ly;                 /* domain lengths (reference: 2*pi) */
  double  dt;                     /* full RK4 step */
  double  viscosity, density, drag;
  int32_t forcing_wavenumber;     /* KolmogorovForcing wave_number (src/inference.py:81) */
  double  forcing_scale;          /* 1.0 in the reference; 0 disables forcing */
  int32_t advection;              /* MLSIM_ADV_* (reference: van Leer) */
  double  lw_dt_scale;            /* 1.0: Lax-Wendroff Courant number uses the full dt */
  int32_t device;                 /* CUDA device ordinal */
  int32_t slab_rank, slab_world;  /* slab_world == 0: plain plan (whole grid on this device).  slab_world >= 1: this
                                   * rank owns global rows [slab_rank*nx/slab_world, (slab_rank+1)*nx/slab_world) of one
                                   * grid decomposed along x; only the mlsim_slab_* / mlsim_cnn_* entry points apply */
  /* Execution tuning (0 = the measured default for the shape).  Results do not depend on these: trajectories are
   * independent, so the block structure only changes how launches overlap. */
  int32_t solver_blocks;          /* 1..8: independent batch blocks whose solver steps run on separate streams
                                   * (default: 2; 4 for bare solver rollouts whose blocks keep >= 2M cells; 1 below 400k cells) */
  int32_t host_blocks;            /* 1..8: equal batch blocks of the H2D | step | D2H pipeline of mlsim_rollout_host
                                   * (default: sized in whole rounds of the persistent conv grid) */
  int32_t separate_first_layer;   /* 0 (default): the network's first layer (2 -> 64) is computed inside the first hidden
                                   * layer's kernel whenever the network allows it, so its activations never reach HBM;
                                   * 1: always a separate first-layer kernel (the round-1 path, kept for A/B and parity) */
  int32_t cluster_projection;     /* 1: single-kernel projection (thread-block cluster per image, half spectrum in
                                   * distributed shared memory); plain square 64 / 128 / 256 grids only.  Measured slower
                                   * than the three-kernel projection at every BASELINE config, hence off by default */
  int32_t step_graph;             /* 1: mlsim_step / mlsim_step_f64 replay a captured CUDA graph of {solver step; correction}
                                   * (the same kernels, arguments and order: results are bit-identical) once the same
                                   * (u, v, p, apply_cnn, scale) has been stepped twice.  Saves host enqueue time only
                                   * (device time unchanged, measured), hence 0 = eager launches by default.  Not used
                                   * while in-step profiling or the finite guard are enabled */
} mlsim_config;

int  mlsim_plan_create (const mlsim_config* cfg, mlsim_plan** out);
int  mlsim_plan_destroy(mlsim_plan* plan);

/* ---- learned correction network: buildModel([CNN block]) of src/models/modelbuilder.py:5-31,
 * src/models/CNN.py:24-40 (3x3, stride 1, zero padding 1, groups 1, ReLU between layers). ----
 * mlsim_cnn_begin declares the layer count; mlsim_cnn_set_layer uploads one layer from HOST memory:
 * w_oihw = Conv2d.weight [cout][cin][3][3], bias [cout], both fp64 (src_dtype 1, the reference's
 * checkpoint dtype, src/train.py:50,56) or fp32 (src_dtype 0).  BatchNorm (bn:true configs) must be
 * folded by the caller.  mlsim_cnn_finalize packs the weights to bf16 operand images on the device. */
int  mlsim_cnn_begin    (mlsim_plan* plan, int32_t nlayers);
int  mlsim_cnn_set_layer(mlsim_plan* plan, int32_t layer, int32_t cin, int32_t cout, int32_t ksize,
                         const void* w_oihw, const void* bias, int32_t src_dtype, int32_t relu_after);
/* ResNet shortcut (reference src/models/ResNet.py:42-50, ``act(x + y)``): after convolution `layer` and before its
 * activation add the block input x = output of layer `src_layer` (-1: the 2-channel network input).
 *   w1x1 == NULL : identity shortcut (hidden layer, same channel count)
 *   w1x1 != NULL : x passes through the block's 1x1 convolution ``conv1`` first: weights [cout][cin_src], bias [cout];
 *                  supported for the network input feeding a hidden layer and for an activation feeding the last
 *                  layer (cout <= 4) -- the two places configs/fullmodels/basicResnet1.json uses it.
 * With relu_after = 1 on that layer the activation is applied to the sum, as the reference does. */
int  mlsim_cnn_set_shortcut(mlsim_plan* plan, int32_t layer, int32_t src_layer, const void* w1x1, const void* b1x1,
                            int32_t src_dtype);
int  mlsim_cnn_finalize (mlsim_plan* plan);

/* Host-only helper (no CUDA call): pack one layer into the bf16 shared-memory operand image the kernels
 * consume.  kind 0 = first layer, 1 = hidden layer, 2 = last layer.  out == NULL returns the byte size. */
int64_t mlsim_cnn_pack_host(int32_t kind, int32_t cin, int32_t cout, const double* w_oihw, const double* bias,
                            void* out, int64_t out_bytes);

/* ---- the hot loop, src/inference.py:100-102:
 *        v, p = step_fn.forward(v, dt, equation=ns2d);  v += model(v)
 * nsteps iterations in place on (u, v).  p_or_null receives the pseudo-pressure of the last
 * projection of the last step.  apply_cnn=0 gives the bare solver step used by
 * src/data_generation.py:206,211,212 and test/sim_test.py:81. */
int  mlsim_step(mlsim_plan* plan, float* u, float* v, float* p_or_null, int32_t nsteps,
                int32_t apply_cnn, double cnn_input_scale, void* stream);

/* The same loop for callers that hold their state in float64, as the reference scripts do
 * (torch.set_default_dtype(torch.float64), src/inference.py:60; RKStepper.from_method(..., dtype=torch.float64), :64):
 * DEVICE pointers to fp64 arrays [batch][nx][ny]; (u_in, v_in) are read, (u_out, v_out) and optionally p_out written
 * (in == out is allowed).  The arithmetic is the fp32 path of mlsim_step -- one fused conversion pass on the way in
 * and one on the way out, on plan-owned staging fields. */
int  mlsim_step_f64(mlsim_plan* plan, const double* u_in, const double* v_in, double* u_out, double* v_out,
                    double* p_out_or_null, int32_t nsteps, int32_t apply_cnn, double cnn_input_scale, void* stream);

/* Same loop through HOST buffers (pinned or pageable): H2D of (u, v), nsteps, D2H of (u, v) and
 * optionally p, synchronised on return.  This is the end-to-end call bench.py times as `e2e`. */
int  mlsim_rollout_host(mlsim_plan* plan, float* host_u, float* host_v, float* host_p_or_null,
                        int32_t nsteps, int32_t apply_cnn, double cnn_input_scale);

/* Explicit block sizes of the H2D | step | D2H pipeline of mlsim_rollout_host: n <= 8 blocks of trajectories whose sizes
 * add up to the batch (sizes == NULL or n == 0: back to the automatic choice).  Results do not depend on it. */
int  mlsim_set_host_blocks(mlsim_plan* plan, const int32_t* sizes, int32_t n);

/* ---- pieces of the step, exposed for parity tests ---- */
/* F(u,v;dt): NavierStokes2DFVMProjection.explicit_terms */
int  mlsim_explicit_terms(mlsim_plan* plan, const float* u, const float* v, float* du, float* dv, void* stream);
/* P(u,v): ...pressure_projection (App. A.4), in place; p_or_null optional */
int  mlsim_project(mlsim_plan* plan, float* u, float* v, float* p_or_null, void* stream);
/* out = irfft2(rfft2(in) * table) on the projection's transform kernels: the spectral filter inside
 * filtered_velocity_field (src/inference.py:76-78, src/data_generation.py:153-155,194-196).  table: HOST fp64
 * [nx][ny/2+1], natural kx order.  Set-up call: synchronises `stream` while the table is uploaded. */
int  mlsim_spectral_multiply(mlsim_plan* plan, const double* table, const float* in, float* out, void* stream);
/* per trajectory: u, v *= max_velocity / max sqrt(u^2 + v^2) (the rescaling step of filtered_velocity_field) */
int  mlsim_normalize_speed(mlsim_plan* plan, float* u, float* v, double max_velocity, void* stream);
/* fdm.divergence(v) (test/sim_test.py:42) */
int  mlsim_divergence(mlsim_plan* plan, const float* u, const float* v, float* div, void* stream);
/* y = model(stack((u,v),1) * input_scale): y is [batch][cout_last][nx][ny] fp32 (NCHW) */
int  mlsim_cnn_forward(mlsim_plan* plan, const float* u, const float* v, float* y_nchw,
                       double cnn_input_scale, void* stream);
/* u += y[:,0]; v += y[:,1] with y = model(...) -- the fused correction-add of src/inference.py:102 */
int  mlsim_cnn_correct(mlsim_plan* plan, float* u, float* v, double cnn_input_scale, void* stream);

/* ---- slab decomposition along x: one large grid over several GPUs (BASELINE.json configs[4]; reference analogue:
 * the same RKStepper.forward / model(v) calls of src/inference.py:100-102 on a grid that does not fit the per-step
 * budget of one device).  The caller owns and exchanges the buffers (NCCL via torch.distributed in slab.py):
 *   field pair   uv[2][batch][ext_rows][ny] fp32, local row i at ext row i + halo  (geometry: mlsim_slab_geometry)
 *   fwd chunks   [world][batch][nxl  ][kywp] float2     bwd chunks  [world][batch][nxl+1][kywp] float2
 * One RK stage = halo exchange of `us` (3 rows from below, 2 from above, periodic) -> mlsim_slab_stage_a (explicit
 * terms + RK combine for rows [-1, nxl), divergence + row FFT -> fwd chunks) -> all-to-all -> mlsim_slab_stage_b
 * (x-transform * inverse Laplacian spectrum * inverse x-transform on this rank's spectral columns -> bwd chunks) ->
 * all-to-all -> mlsim_slab_stage_c (row inverse FFT + gradient subtraction, in place on `uv`).
 * mlsim_slab_cnn_correct needs 7 current halo rows per interior side (zero padding at the global edges). */
/* out8 = {nxl, halo, ext_rows, kyw, kywp, kyv (this rank's valid spectral columns), outer radix r0, cnn halo} */
int  mlsim_slab_geometry(mlsim_plan* plan, int64_t* out8);
int  mlsim_slab_stage_a(mlsim_plan* plan, int32_t stage, const float* us, const float* u0, float* acc, float* out,
                        void* spec_send, void* stream);
int  mlsim_slab_divergence_rfft(mlsim_plan* plan, const float* uv, void* spec_send, void* stream);   /* P alone: first half */
int  mlsim_slab_stage_b(mlsim_plan* plan, const void* spec_recv, void* spec_send2, void* stream);
int  mlsim_slab_stage_c(mlsim_plan* plan, const void* spec_recv2, float* uv, float* p_or_null, void* stream);
int  mlsim_slab_cnn_correct(mlsim_plan* plan, float* uv, double cnn_input_scale, void* stream);
/* Peer-memory (NVLink P2P) spectral exchange.  fwd_recv[r] / bwd_recv[r], r < slab_world: device address, valid in THIS
 * process, of rank r's forward / backward receive buffer (CUDA IPC or symmetric memory; slab.py uses
 * torch.distributed._symmetric_memory).  Afterwards mlsim_slab_stage_a with spec_send == NULL stores column chunk r
 * straight into rank r's receive buffer, mlsim_slab_stage_b with spec_send2 == NULL does the same for the backward
 * chunks, and the caller replaces each all-to-all by a cross-rank barrier.  NULL arguments switch back. */
int  mlsim_slab_set_peers(mlsim_plan* plan, const uint64_t* fwd_recv, const uint64_t* bwd_recv);
/* Peer-memory halo exchange: copy this rank's top `lo` rows into the upper neighbour's low halo rows and its bottom `hi`
 * rows into the lower neighbour's high halo rows.  up_uv / dn_uv: the neighbours' field pair of the same role as mapped
 * into this process (NULL: skip that side -- the global edges of the non-periodic CNN halo).  Follow with a barrier. */
int  mlsim_slab_halo_push(mlsim_plan* plan, const float* uv, float* up_uv, float* dn_uv, int32_t lo, int32_t hi, void* stream);

/* ---- data generation, reference src/data_generation.py:91-111,193-213 ----
 * mlsim_downsample_staggered: downsample_staggered_velocity(fine_grid, coarse_grid, (u, v)) -- u keeps the faces
 * factor-1::factor along x and is block-averaged over y, v the faces along y averaged over x.  Inputs
 * [batch][nx][ny], outputs [batch][nx/factor][ny/factor], fp32 device pointers; needs no plan.
 * mlsim_datagen_pair: one stored training pair, entirely on the device and stream-ordered:
 *   n_between fine steps (:205-206); (cu, cv) = downsample(u, v) (:209); n_inner fine steps (:210-211);
 *   one coarse step of (cu, cv) with the coarse plan's dt (:212).  On return (u, v) is the `_f` entry and
 *   (cu, cv) the `_c` entry of the sample (:219-220). */
int  mlsim_downsample_staggered(const float* u, const float* v, float* cu, float* cv, int32_t batch, int32_t nx,
                                int32_t ny, int32_t factor, void* stream);
int  mlsim_datagen_pair(mlsim_plan* fine, mlsim_plan* coarse, float* u, float* v, float* cu, float* cv,
                        int32_t n_between, int32_t n_inner, void* stream);

/* ---- divergence guard (reference test/sim_test.py:82-87 tests isnan on the host after every step) ----
 * With every > 0, after each `every`-th step of mlsim_step / mlsim_rollout_host a small kernel scans (u, v) and
 * publishes, in a pinned host word, the index of the first step (counted from this call, starting at 1) after which
 * a non-finite value was present.  No synchronisation is added to the step path.  every = 0 disables the guard. */
int  mlsim_finite_check_enable(mlsim_plan* plan, int32_t every);
/* Non-blocking: *first_bad_step = what the device has published so far (-1: every check so far passed);
 * *steps_issued = steps enqueued since mlsim_finite_check_enable.  Synchronise the stream first for a final answer. */
int  mlsim_finite_check_read(mlsim_plan* plan, int64_t* first_bad_step, int64_t* steps_issued);

/* ---- introspection ---- */
/* number of kernel launches one mlsim_step(nsteps=1) issues (solver only, or solver + cnn) */
int  mlsim_launches_per_step(mlsim_plan* plan, int32_t apply_cnn);
/* number of cudaGraphLaunch calls mlsim_step / mlsim_step_f64 have issued so far (mlsim_config.step_graph); -1: null plan */
int64_t mlsim_step_graph_launches(mlsim_plan* plan);
/* bytes of device workspace owned by the plan */
int64_t mlsim_workspace_bytes(mlsim_plan* plan);
/* time one kernel class in isolation with CUDA events on `stream`: which = 0 explicit terms (stage 2
 * form), 1 div+row FFT, 2 column solve, 3 row iFFT+gradient, 4 first conv, 5 one mid conv layer,
 * 6 last conv + correction add, 7 single-kernel cluster projection (small square grids), 8 first conv layer fused into
 * the first hidden layer.  Runs `iters` launches after `warmup`, returns mean ms in *ms_out. */
int  mlsim_time_kernel(mlsim_plan* plan, int32_t which, int32_t warmup, int32_t iters, float* ms_out, void* stream);

/* Developer aid: mean per-CTA cycle counters of the last conv kernel run by mlsim_time_kernel(5|6):
 * out8 = {MMA-warp total, waiting for drained accumulators, waiting for staged input rows, issuing,
 * input rows, epilogue-warp total, epilogue waiting for finished accumulators, unused}.  Returns #CTAs. */
int  mlsim_conv_debug(mlsim_plan* plan, double* out8);
/* Same for the first-layer group of the fused kernel (mlsim_time_kernel(8)): out8 = {group total, waiting for a released
 * stage, waiting for its accumulator, draining (TMEM -> ReLU -> bf16 -> stage), building the next A tile (incl. the
 * row loads), rows, group barriers, unused}. */
int  mlsim_conv_debug_first(mlsim_plan* plan, double* out8);

/* In-step profiling: bracket every launch of one kernel class (numbering as in mlsim_time_kernel; -1 = off)
 * with CUDA events on the launching stream, for up to max_samples launches; mlsim_profile_read synchronises
 * on the recorded events, returns the summed device time and the number of launches, and resets the count. */
int  mlsim_profile_enable(mlsim_plan* plan, int32_t kernel_class, int32_t max_samples);
int  mlsim_profile_read(mlsim_plan* plan, float* total_ms, int32_t* count);

const char* mlsim_last_error(void);
const char* mlsim_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MLSIM_H */

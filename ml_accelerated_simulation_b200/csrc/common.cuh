// Shared host/device helpers for the mlsim sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mlsim.h"

namespace mlsim {

void set_error(const std::string& msg);

#define MLSIM_CUDA_CHECK(expr)                                                          \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::mlsim::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
      return MLSIM_ECUDA;                                                               \
    }                                                                                   \
  } while (0)

#define MLSIM_REQUIRE(cond, msg)                                                        \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      ::mlsim::set_error(std::string(msg) + " [" #cond "]");                            \
      return MLSIM_EINVAL;                                                              \
    }                                                                                   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Solver parameters handed to the kernels by value.
struct SolverParams {
  int nx, ny, batch;
  int nyc;        // ny/2 + 1 spectral columns
  int nycp;       // padded pitch of the spectral buffer (float2 units)
  float inv_hx, inv_hy;       // 1/h
  float nu_hx2, nu_hy2;       // (viscosity/density)/h^2
  float dt_hx, dt_hy;         // lw_dt_scale*dt/h  (Courant factor)
  float drag;
  int advection;
  long long img_stride;       // elements between consecutive images of a field (plain: nx*ny; slab: ext_rows*ny)
  // slab decomposition along x (mlsim_config.slab_world >= 1): nx above is the LOCAL row count, field pointers
  // address local row 0 of halo-extended buffers (rows [-halo, nx+halo) are valid memory, no periodic wrap in x)
  int halo;
  int kyw, kywp;              // spectral columns per rank chunk and its padded pitch (float2 units)
  int raw;                    // 1: K2 transforms the field `u` itself (no divergence), K4 only writes the inverse
                              //    transform to `pout` -- together with K3 and a caller-supplied spectral table this is
                              //    out = irfft2(rfft2(in) * table) (initial-condition filter, mlsim_spectral_multiply)
};

// Slab mode: where spectral column chunk r lives.  With NCCL exchanges every entry points into one local send / recv
// buffer (chunk r at r * chunk elements); with peer-memory exchanges entry r is rank r's receive buffer mapped into this
// process (NVLink P2P), already offset to this rank's slot, so the producing kernel stores straight into the consumer's
// memory and the all-to-all disappears.
#define MLSIM_MAX_SLAB_RANKS 8
struct SpecPeers {
  float2* p[MLSIM_MAX_SLAB_RANKS];
};
#ifdef __CUDACC__
__device__ __forceinline__ float2* spec_peer(const SpecPeers& s, int r) {
  float2* q = s.p[0];
#pragma unroll
  for (int i = 1; i < MLSIM_MAX_SLAB_RANKS; ++i)
    if (r == i) q = s.p[i];
  return q;
}
#endif

// Stage coefficients of the RK4-with-projection scheme.
//   mode 0: du,dv = F                       (test hook)
//   mode 1: acc = u0 + ca*k ; ustar = u0 + cs*k            (stage 1: u_s == u0)
//   mode 2: acc = acc + ca*k ; ustar = u0 + cs*k           (stages 2,3)
//   mode 3: ustar = acc + cs*k                              (stage 4)
struct StageArgs {
  const float* us; const float* vs;   // state the explicit terms are evaluated on
  const float* u0; const float* v0;   // state at the start of the step
  float* accu; float* accv;           // running sum u0 + dt*sum(b_j k_j)
  float* outu; float* outv;           // ustar (or du,dv in mode 0)
  float ca, cs;
  int mode;
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the step is launched with the "programmatic stream serialization"
// attribute, releases its dependents at once (pdl_launch_dependents) and waits for the complete, flushed result of
// its predecessor (pdl_wait) before touching any buffer the step writes.  The next kernel's CTAs are therefore
// already resident (prologue done: indices, barrier init, TMEM allocation, weight image load) when the previous
// grid drains, instead of paying drain + launch latency between each of the 23 dependent kernels of a step.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (kernel, device), so that a
// second plan on another GPU of the same process gets it too.  `kernel` may contain template commas: pass it in parentheses.
#define MLSIM_MAX_DEVICES 64
#define MLSIM_SET_SMEM_ONCE(kernel, bytes)                                                                      \
  do {                                                                                                          \
    static bool done_[MLSIM_MAX_DEVICES] = {};                                                                  \
    int dev_ = 0;                                                                                               \
    MLSIM_CUDA_CHECK(cudaGetDevice(&dev_));                                                                     \
    const bool in_ = dev_ >= 0 && dev_ < MLSIM_MAX_DEVICES;                                                     \
    if (!in_ || !done_[dev_]) {                                                                                 \
      MLSIM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      if (in_) done_[dev_] = true;                                                                              \
    }                                                                                                           \
  } while (0)

static inline int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
static inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

}  // namespace mlsim

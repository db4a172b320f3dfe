// Launchers of the solver kernels (definitions in solver_kernels.cu).
#pragma once
#include "common.cuh"

namespace mlsim {
int launch_explicit(const StageArgs& a, const SolverParams& p, const float* forcing, cudaStream_t st);
int launch_divergence(const float* u, const float* v, float* d, const SolverParams& p, cudaStream_t st);
int launch_div_rfft(const float* u, const float* v, float2* S, const float2* tw_y, const SolverParams& p,
                    cudaStream_t st);
int launch_col_solve(float2* S, const float* inv_eig, const float2* tw_x, const SolverParams& p, cudaStream_t st);
int launch_irfft_grad(const float2* S, const float* uin, const float* vin, float* uout, float* vout, float* pout,
                      const float2* tw_y, const SolverParams& p, cudaStream_t st);
// single-kernel projection (thread-block cluster per image) for small square grids
bool project_cluster_supported(const SolverParams& p);
int launch_project_cluster(float* u, float* v, float* pout, const float* inv_eig, const float2* tw,
                           const SolverParams& p, cudaStream_t st);
}  // namespace mlsim

// Launchers of the solver kernels (definitions in solver_kernels.cu).
#pragma once
#include "common.cuh"

namespace mlsim {
int launch_explicit(const StageArgs& a, const SolverParams& p, const float* forcing, cudaStream_t st);
int launch_divergence(const float* u, const float* v, float* d, const SolverParams& p, cudaStream_t st);
int launch_div_rfft(const float* u, const float* v, float2* S, const float2* tw_y, const SolverParams& p,
                    cudaStream_t st, const SpecPeers* peers = nullptr, int world = 1);
int launch_col_solve(float2* S, const float* inv_eig, const float2* tw_x, const SolverParams& p, int n, int images,
                     int eig_images, cudaStream_t st);
int launch_col_solve_split(float2* S, float2* T, const float* inv_eig_perm, const float2* twx_full, const float2* tw_n2,
                           const SolverParams& p, int r0, cudaStream_t st);
int launch_col_solve_slab(const float2* recv, float2* send2, float2* T, const float* inv_eig_perm, const float2* twx_full,
                          const float2* tw_n2, const SolverParams& p, int nxg, int world, int kyv, int r0, cudaStream_t st,
                          const SpecPeers* peers = nullptr);
int launch_irfft_grad(const float2* S, const float* uin, const float* vin, float* uout, float* vout, float* pout,
                      const float2* tw_y, const SolverParams& p, cudaStream_t st, int world = 1);
// single-kernel projection (thread-block cluster per image) for small square grids
bool project_cluster_supported(const SolverParams& p);
int launch_project_cluster(float* u, float* v, float* pout, const float* inv_eig, const float2* tw,
                           const SolverParams& p, cudaStream_t st);
// datagen_kernels.cu
int launch_normalize_speed(float* u, float* v, unsigned int* scratch, int batch, int n_per_img, float max_velocity,
                           cudaStream_t st);
int launch_halo_push(const float* uv, float* up_uv, float* dn_uv, int images, int ext_rows, int ny, int nxl, int halo,
                     int lo, int hi, cudaStream_t st);
int launch_finite_check(const float* u, const float* v, size_t n, unsigned long long step, unsigned long long* flag,
                        cudaStream_t st);
int launch_cast_in(const double* a, const double* b, float* oa, float* ob, size_t n, cudaStream_t st);
int launch_cast_out(const float* a, const float* b, const float* c, double* oa, double* ob, double* oc, size_t n,
                    cudaStream_t st);
int launch_downsample_staggered(const float* u, const float* v, float* cu, float* cv, int batch, int nx, int ny, int f,
                                cudaStream_t st);
}  // namespace mlsim

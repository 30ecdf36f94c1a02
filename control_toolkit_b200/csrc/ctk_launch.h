// ctk_launch.h -- host-callable launch wrappers; each is defined in the translation unit that owns the kernels so the
// library builds as parallel nvcc jobs.  All launches are asynchronous on `st` and return cudaGetLastError().
#pragma once
#include <cuda_runtime.h>

#include "ctk_args.cuh"

#include <cstdlib>
#include <utility>

namespace ctk {

// Launch with programmatic stream serialization (see pdl_wait in ctk_device.cuh).  CTK_NO_PDL=1 falls back to plain launches.
inline bool pdl_enabled() {
  static const bool on = std::getenv("CTK_NO_PDL") == nullptr;
  return on;
}
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

// pred: 0 ODE, 1 MLP (SIMT engine); kind: cost class; log: write SoA trajectory logs
cudaError_t launch_mppi_rollout(int pred, int kind, bool log, int nblocks, int block, size_t smem, cudaStream_t st, const MppiArgs& a);
int mppi_max_block_threads(int pred);
// K1 for the ODE predictor with intermediate_steps == 1 (scaled variables, ILP rollouts per thread)
cudaError_t launch_mppi_ode(int kind, bool log, int period_t, int ilp, int grid, int block, size_t smem, cudaStream_t st, const MppiOdeArgs& a,
                            const char** name = nullptr);
int mppi_ode_max_block(int ilp, bool log);
// several clients' ticks in one launch: grid (grid, nclients); Philox noise, logging off, ILP 1 (ctk_batch.cu)
cudaError_t launch_mppi_ode_batch(int kind, int period_t, int grid, int nclients, int block, size_t smem, cudaStream_t st, const MppiOdeArgs& a,
                                  const MppiBatch& b);
size_t mppi_ode_smem_bytes(int H, int period, int n_ind, int ilp, int block);
cudaError_t launch_mppi_combine(const float* in, int cnt, int n_ind, float neg_inv_lbd, float* record_out,
                                const MppiFinalize& fin, cudaStream_t st);
cudaError_t launch_exchange_barrier(const MppiFuse& f, size_t bar_off, cudaStream_t st);
cudaError_t launch_transpose(const float* in, float* out, int R, int C, cudaStream_t st);
// top-M logging: gather the rows of the rollouts named by m sorted K4 keys out of an SoA [R][N] or reference-layout [N][R] log
cudaError_t launch_log_gather(const uint64_t* keys, int m, int off, const float* J, size_t N, const float* src, int R, int soa, float* out,
                              float* J_out, int32_t* idx_out, cudaStream_t st);
size_t mppi_pred_smem_floats(int pred, const MlpDev& m);
// environments beyond the CartPole (ctk_env.cu): MPPI = rollout + update launches, CEM = rollout, K4 levels, refit
cudaError_t launch_env_mppi(int env, bool log, const EnvMppiArgs& a, cudaStream_t st);
cudaError_t launch_env_cem_rollout(int env, bool log, const EnvCemArgs& a, cudaStream_t st);
cudaError_t launch_env_cem_refit(const EnvCemRefitArgs& a, cudaStream_t st);
void env_dims(int env, int* ns, int* nu);
// recurrent predictor (ctk_gru.cu)
cudaError_t launch_mppi_rollout_gru(int kind, bool log, int nblocks, int block, size_t smem, cudaStream_t st, const MppiArgs& a);
cudaError_t launch_cem_rollout_gru(int kind, bool log, int nblocks, size_t smem, cudaStream_t st, const CemArgs& a);
cudaError_t launch_single_rollout_gru(const float* s0, const float* Q, int H, const DevConsts* kc, const MlpDev& mlp, const float* u_prev,
                                      float* traj, float* summed, cudaStream_t st);
cudaError_t launch_gru_update(const S0& s0, const float* u_nom, const MlpDev& mlp, float* rnn_h, cudaStream_t st);

cudaError_t launch_cem_rollout(int pred, int kind, bool log, int nblocks, size_t smem, cudaStream_t st, const CemArgs& a);
int cem_rollout_grid(int pred, int N, int num_sms);
cudaError_t launch_cem_ode(int kind, bool log, int nblocks, size_t smem, cudaStream_t st, const CemOdeArgs& a);
cudaError_t launch_cem_tick(int kind, bool log, int nblocks, size_t smem, cudaStream_t st, const CemTickArgs& a);
int cem_tick_rollouts_per_block();
int cem_tick_blocks_per_sm(int kind, bool log, size_t smem);
cudaError_t launch_topk_level(const float* cost, const uint64_t* keys_in, int n, int off, int k, uint64_t* out, cudaStream_t st);
cudaError_t launch_cem_refit(const CemRefitArgs& a, cudaStream_t st);

// coef: coefficient-form adjoint (rpgd_grad_coef_kernel, tape of 12 floats per step) instead of the direct form (8 floats per step)
// fused_select != null (coef only, one block): the select / resample step (K8) runs at the end of the same launch
cudaError_t launch_rpgd_grad(int kind, bool log, bool coef, int nblocks, int block, size_t smem, cudaStream_t st, const RpgdGradArgs& a,
                             const RpgdSelectArgs* fused_select = nullptr);
cudaError_t launch_rpgd_select(const RpgdSelectArgs& a, cudaStream_t st);
cudaError_t launch_rpgd_init(const RpgdSelectArgs& a, cudaStream_t st);
cudaError_t launch_gradcem_sample(const GradCemSampleArgs& a, cudaStream_t st);
cudaError_t launch_gradcem_refit(const GradCemRefitArgs& a, cudaStream_t st);

cudaError_t launch_single_rollout(int pred, const float* s0, const float* Q, int H, const DevConsts* kc, const MlpDev& mlp, const float* u_prev, float* traj, float* summed, cudaStream_t st);
cudaError_t launch_fma_peak(float* out, int blocks, int threads, int iters, cudaStream_t st);
cudaError_t launch_fp32_micro(int variant, float* out, int blocks, int threads, int iters, const float* seed, cudaStream_t st);
cudaError_t launch_philox_fill(const NoiseSrc& ns, float* out, size_t n, cudaStream_t st);
cudaError_t launch_philox_export(const NoiseSrc& ns, size_t row0, float* out, size_t n, cudaStream_t st);

}  // namespace ctk

// ctk_gru.cu -- translation unit owning the instantiations of the rollout kernels for the recurrent (GRU) predictor
// (GruSimtPred, ctk_predictor.cuh) and its state hook, predictor.update (reference optimizer_mppi.py:192,195-197).
#include "ctk_kernels_mppi.cuh"
#include "ctk_kernels_cem.cuh"
#include "ctk_launch.h"

namespace ctk {

template <int KIND, bool LOG>
static cudaError_t launch_mppi_gru_t(int nblocks, int block, size_t smem, cudaStream_t st, const MppiArgs& a) {
  auto k = mppi_rollout_kernel<GruSimtPred, KIND, LOG>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<nblocks, block, smem, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mppi_rollout_gru(int kind, bool log, int nblocks, int block, size_t smem, cudaStream_t st, const MppiArgs& a) {
  if (kind == 0) return log ? launch_mppi_gru_t<0, true>(nblocks, block, smem, st, a) : launch_mppi_gru_t<0, false>(nblocks, block, smem, st, a);
  return log ? launch_mppi_gru_t<1, true>(nblocks, block, smem, st, a) : launch_mppi_gru_t<1, false>(nblocks, block, smem, st, a);
}

template <int KIND, bool LOG>
static cudaError_t launch_cem_gru_t(int nblocks, size_t smem, cudaStream_t st, const CemArgs& a) {
  auto k = cem_rollout_kernel<GruSimtPred, KIND, LOG>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(k, dim3(nblocks), dim3(GruSimtPred::kCemThreads), smem, st, a);
}
cudaError_t launch_cem_rollout_gru(int kind, bool log, int nblocks, size_t smem, cudaStream_t st, const CemArgs& a) {
  if (kind == 0) return log ? launch_cem_gru_t<0, true>(nblocks, smem, st, a) : launch_cem_gru_t<0, false>(nblocks, smem, st, a);
  return log ? launch_cem_gru_t<1, true>(nblocks, smem, st, a) : launch_cem_gru_t<1, false>(nblocks, smem, st, a);
}

// predictor.update(s, Q0) of the recurrent predictor (reference optimizer_mppi.py:192,195-197): the saved hidden state advances one
// step with the measured state and the first control of the tick's updated nominal sequence; row 1 keeps the state the tick's
// rollouts started from (nominal rollout, :199-202).  One block; thread 0 walks the cell.
__global__ void __launch_bounds__(GruSimtPred::kMaxThreads) gru_update_kernel(const S0 s0, const float* u_nom, const MlpDev mlp, float* rnn_h) {
  extern __shared__ float smem[];
  GruSimtPred pred(nullptr, mlp, smem);
  __syncthreads();
  if (threadIdx.x != 0) return;
  State z;
  z.th = s0.ld(0); z.om = s0.ld(1); z.c = s0.ld(2); z.s = s0.ld(3); z.x = s0.ld(4); z.v = s0.ld(5);
  float omc = 0.0f;
  pred.begin_rollout();
  pred.step(z, u_nom[0], omc);
  const int hid = mlp.hidden;
  for (int l = 0; l < 2; ++l)
    for (int j = 0; j < hid; ++j) {
      rnn_h[2 * hid + l * hid + j] = rnn_h[l * hid + j];
      rnn_h[l * hid + j] = pred.hs[(l * GruSimtPred::HMAX + j) * GruSimtPred::kMaxThreads];
    }
}
cudaError_t launch_gru_update(const S0& s0, const float* u_nom, const MlpDev& mlp, float* rnn_h, cudaStream_t st) {
  const size_t smem = sizeof(float) * GruSimtPred::smem_floats(mlp);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gru_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  gru_update_kernel<<<1, 32, smem, st>>>(s0, u_nom, mlp, rnn_h);
  return cudaGetLastError();
}


struct GruPrevPred : GruSimtPred {  // the nominal rollout starts from the hidden state BEFORE the tick's predictor.update
  __device__ __forceinline__ GruPrevPred(const DevConsts* kc, const MlpDev& m, float* sm) : GruSimtPred(kc, m, sm, 1) {}
};
// nominal rollout of one control sequence (reference optimizer_mppi.py:199-202 predict_optimal_trajectory)
__global__ void single_rollout_gru_kernel(const float* s0, const float* Q, int H, const DevConsts* kc, MlpDev mlp, const float* u_prev,
                                          float* traj, float* summed) {
  extern __shared__ float smem[];
  GruPrevPred pred(kc, mlp, smem);
  const CostC cost = kc->cost;
  __syncthreads();
  if (threadIdx.x != 0) return;
  State z;
  z.th = s0[0]; z.om = s0[1]; z.c = s0[2]; z.s = s0[3]; z.x = s0[4]; z.v = s0[5];
  float omc = 1.0f - cosf(z.th), ul = u_prev[0], sum = 0.0f;
  pred.begin_rollout();
  for (int t = 0; t <= H; ++t) {
    float* p = traj + t * 6;
    p[0] = z.th; p[1] = z.om; p[2] = z.c; p[3] = z.s; p[4] = z.x; p[5] = z.v;
    if (t == H) break;
    const float u = Q[t];
    sum += stage_cost_dyn(cost.kind, z, omc, u, ul, cost);
    pred.step(z, u, omc);
    ul = u;
  }
  summed[0] = (sum - cost.shift) * (float)(H + 1);
}
cudaError_t launch_single_rollout_gru(const float* s0, const float* Q, int H, const DevConsts* kc, const MlpDev& mlp, const float* u_prev,
                                      float* traj, float* summed, cudaStream_t st) {
  const size_t smem = sizeof(float) * GruSimtPred::smem_floats(mlp);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(single_rollout_gru_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  single_rollout_gru_kernel<<<1, 32, smem, st>>>(s0, Q, H, kc, mlp, u_prev, traj, summed);
  return cudaGetLastError();
}

}  // namespace ctk

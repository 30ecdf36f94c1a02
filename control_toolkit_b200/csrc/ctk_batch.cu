// ctk_batch.cu -- translation unit owning the multi-client MPPI tick (mppi_ode_batch_kernel: SURVEY 8f.4, several remote clients'
// states in ONE launch behind the serving edge; reference controller_server/controller_server.py:55-86 steps one client per request).
#include "ctk_kernels_mppi_ode.cuh"
#include "ctk_launch.h"

namespace ctk {

template <int KIND, int PERIOD>
static cudaError_t launch_batch_t(int grid, int nclients, int block, size_t smem, cudaStream_t st, const MppiOdeArgs& a, const MppiBatch& b) {
  auto k = mppi_ode_batch_kernel<KIND, PERIOD, 1024>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<dim3(grid, nclients), dim3(block), smem, st>>>(a, b);
  return cudaGetLastError();
}

// in-kernel Philox noise, logging off, one rollout per thread: the instantiation a single client's handle runs at these sizes
cudaError_t launch_mppi_ode_batch(int kind, int period_t, int grid, int nclients, int block, size_t smem, cudaStream_t st, const MppiOdeArgs& a,
                                  const MppiBatch& b) {
  if (kind == 0) return period_t == 10 ? launch_batch_t<0, 10>(grid, nclients, block, smem, st, a, b) : launch_batch_t<0, 0>(grid, nclients, block, smem, st, a, b);
  return period_t == 10 ? launch_batch_t<1, 10>(grid, nclients, block, smem, st, a, b) : launch_batch_t<1, 0>(grid, nclients, block, smem, st, a, b);
}

}  // namespace ctk

// ctk_env.cu -- translation unit owning the general-environment kernels (ctk_kernels_env.cuh; SURVEY 8f.3: environments beyond the
// CartPole through the functor registry).  env: 1 = Dubins car (3 states, 2 control inputs).
#include "ctk_kernels_env.cuh"
#include "ctk_launch.h"

namespace ctk {

cudaError_t launch_env_mppi(int env, bool log, const EnvMppiArgs& a, cudaStream_t st) {
  if (env != 1) return cudaErrorInvalidValue;
  const int nb = (a.N + 127) / 128;
  if (log) env_mppi_rollout_kernel<DubinsEnv, true><<<nb, 128, 0, st>>>(a);
  else env_mppi_rollout_kernel<DubinsEnv, false><<<nb, 128, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  env_mppi_update_kernel<DubinsEnv::NU><<<1, 1024, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_env_cem_rollout(int env, bool log, const EnvCemArgs& a, cudaStream_t st) {
  if (env != 1) return cudaErrorInvalidValue;
  const int nb = (a.N + 127) / 128;
  if (log) env_cem_rollout_kernel<DubinsEnv, true><<<nb, 128, 0, st>>>(a);
  else env_cem_rollout_kernel<DubinsEnv, false><<<nb, 128, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_env_cem_refit(const EnvCemRefitArgs& a, cudaStream_t st) {
  env_cem_refit_kernel<<<1, TOPK_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}
void env_dims(int env, int* ns, int* nu) {
  *ns = env == 1 ? DubinsEnv::NS : 6;
  *nu = env == 1 ? DubinsEnv::NU : 1;
}

}  // namespace ctk

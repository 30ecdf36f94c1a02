// ctk_kernels_cem.cuh -- K3 fused sample -> rollout -> cost for CEM, K5 elite merge + refit (+ post-loop shift).
// Replaces reference optimizer_cem_tf.py:54-80 (update_distribution) and :99-102 (post-loop) for one tick.
#pragma once
#include "ctk_device.cuh"
#include "ctk_predictor.cuh"
#include "ctk_topk.cuh"
#include "ctk_ode_scaled.cuh"

namespace ctk {

// Q[n,t] = clip(mu_t + z * sd_t): optimizer_cem_tf.py:64-66 (tf.multiply then add, separately rounded)
__device__ __forceinline__ float cem_sample(float mu, float sd, float z, float lo, float hi) {
  return fminf(fmaxf(__fadd_rn(mu, __fmul_rn(z, sd)), lo), hi);
}

template <class Pred, int KIND, bool LOG>
__global__ void __launch_bounds__(128) cem_rollout_kernel(const CemArgs a) {
  extern __shared__ float smem[];
  float* sh_mu = smem;         // [H]
  float* sh_sd = smem + a.H;   // [H]
  Pred pred(a.kc, a.mlp, smem + 2 * a.H);  // constants: independent of the previous kernel, loaded while it drains
  const CostC cost = load_cost(a.kc);
  pdl_wait();
  pdl_trigger();
  for (int t = threadIdx.x; t < a.H; t += blockDim.x) {
    sh_mu[t] = a.mu[t];
    sh_sd[t] = a.sd[t];
  }
  __syncthreads();

  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < a.N;
  if (!active && !Pred::kCooperative) return;
  const uint32_t ng = (uint32_t)(a.off + (active ? n : 0));

  State z;
  z.th = a.s0.ld(0); z.om = a.s0.ld(1); z.c = a.s0.ld(2); z.s = a.s0.ld(3); z.x = a.s0.ld(4); z.v = a.s0.ld(5);
  float omc = 1.0f - cosf(z.th);
  float u_last = a.u_prev[0];
  float jsum = 0.0f;
  for (int t0 = 0; t0 < a.H; t0 += 4) {
    float zz[4];
    noise4(a.noise, ng, (uint32_t)(t0 >> 2), zz);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = t0 + q;
      if (t < a.H) {
        const float u = cem_sample(sh_mu[t], sh_sd[t], zz[q], a.lo, a.hi);
        if (LOG && active) {
          float* p = a.log_traj_soa + (size_t)t * 6 * a.N + n;
          p[0] = z.th; p[a.N] = z.om; p[2 * a.N] = z.c; p[3 * a.N] = z.s; p[4 * (size_t)a.N] = z.x; p[5 * (size_t)a.N] = z.v;
          a.log_Q_soa[(size_t)t * a.N + n] = u;
        }
        jsum += stage_cost<KIND>(z, omc, u, u_last, cost);
        pred.step(z, u, omc);
        u_last = u;
      }
    }
  }
  if (!active) return;
  if (LOG) {
    float* p = a.log_traj_soa + (size_t)a.H * 6 * a.N + n;
    p[0] = z.th; p[a.N] = z.om; p[2 * a.N] = z.c; p[3 * a.N] = z.s; p[4 * (size_t)a.N] = z.x; p[5 * (size_t)a.N] = z.v;
  }
  a.J[n] = (jsum + terminal_cost(z, cost)) - cost.shift;
}

// One block of 1024 threads: merge candidates -> global top-k (bitonic), regenerate elite Q, refit mu / sd.
// K3s: the same sample -> rollout -> cost pass for the ODE predictor in the scaled state variables of K1 (12-instruction Euler
// step, cost with the control terms merged into u (kA u + kB u_prev) and the u_prev^2 terms telescoped; constants are kernel
// parameters -> uniform registers).  The Philox block of the NEXT four steps is generated next to the current four steps'
// dependent chain, so the draw latency never sits on it.
constexpr int kCemOdeTopkThreads = 256;  // block size when the block-level top-k is fused in (CemOdeArgs::cand_out)
template <int KIND, bool LOG>
__global__ void __launch_bounds__(kCemOdeTopkThreads) cem_ode_kernel(const CemOdeArgs a) {
  extern __shared__ float smem[];
  float* sh_mu = smem;         // [H]
  float* sh_sd = smem + a.H;   // [H]
  __shared__ uint64_t sh_keys[kCemOdeTopkThreads];
  const OdeHot& k = a.k;
  pdl_wait();
  pdl_trigger();
  for (int t = threadIdx.x; t < a.H; t += blockDim.x) {
    sh_mu[t] = a.mu[t];
    sh_sd[t] = a.sd[t];
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < a.N;
  if (!active && a.cand_out == nullptr) return;
  uint64_t key = KEY_MAX;
  if (active) {
  const uint32_t ng = (uint32_t)(a.off + n);
  const float s0v[6] = {a.s0.ld(0), a.s0.ld(1), a.s0.ld(2), a.s0.ld(3), a.s0.ld(4), a.s0.ld(5)};
  ScaledState r;
  scaled_from_state(s0v, k, r);
  const float u_prev = a.u_prev[0];
  float ul = u_prev;
  float acc = (k.k_ccrc * u_prev) * u_prev;  // telescoped control-change cost: + k_ccrc u_{-1}^2 here, - k_ccrc u_{H-1}^2 at the end
  float zn[4];
  noise4(a.noise, ng, 0u, zn);
  for (int t0 = 0; t0 < a.H; t0 += 4) {
    const float z0 = zn[0], z1 = zn[1], z2 = zn[2], z3 = zn[3];
    if (t0 + 4 < a.H) noise4(a.noise, ng, (uint32_t)((t0 >> 2) + 1), zn);
    const float zz[4] = {z0, z1, z2, z3};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = t0 + q;
      if (t < a.H) {
        const float u = cem_sample(sh_mu[t], sh_sd[t], zz[q], k.lo, k.hi);
        if (LOG) {
          float st[6];
          scaled_to_state(r, k, st);
          float* p = a.log_traj_soa + (size_t)t * 6 * a.N + n;
          p[0] = st[0]; p[a.N] = st[1]; p[2 * a.N] = st[2]; p[3 * a.N] = st[3]; p[4 * (size_t)a.N] = st[4]; p[5 * (size_t)a.N] = st[5];
          a.log_Q_soa[(size_t)t * a.N + n] = u;
        }
        acc = stage_cost_scaled<KIND>(acc, r, u, ul, 0.0f, k);  // kC == 0 for CEM (no MPPI correction)
        ode_step_scaled(r, u, k);
        ul = u;
      }
    }
  }
  if (LOG) {
    float st[6];
    scaled_to_state(r, k, st);
    float* p = a.log_traj_soa + (size_t)a.H * 6 * a.N + n;
    p[0] = st[0]; p[a.N] = st[1]; p[2 * a.N] = st[2]; p[3 * a.N] = st[3]; p[4 * (size_t)a.N] = st[4]; p[5 * (size_t)a.N] = st[5];
  }
  const float J = finish_cost_scaled(acc, r, ul, k);
  a.J[n] = J;
  key = make_key(J, ng);
  }
  if (a.cand_out != nullptr) {  // block-level top-k (K4 level 0): blockDim.x == kCemOdeTopkThreads
    key = block_bitonic_sort(key, sh_keys, kCemOdeTopkThreads);
    if ((int)threadIdx.x < a.kk) a.cand_out[(size_t)blockIdx.x * a.kk + threadIdx.x] = key;
  }
}

__global__ void __launch_bounds__(TOPK_THREADS) cem_refit_kernel(const CemRefitArgs a) {
  __shared__ uint64_t sh[TOPK_THREADS];
  __shared__ uint32_t sh_elite[TOPK_THREADS];
  pdl_wait();
  pdl_trigger();
  uint64_t key = (threadIdx.x < a.cnt) ? a.cand[threadIdx.x] : KEY_MAX;
  int n_sort = 32;
  while (n_sort < a.cnt) n_sort <<= 1;
  key = block_bitonic_sort(key, sh, n_sort);
  if (threadIdx.x < a.k) {
    sh_elite[threadIdx.x] = (uint32_t)(key & 0xffffffffu);
    if (a.elite_idx_out != nullptr) a.elite_idx_out[threadIdx.x] = (int32_t)(key & 0xffffffffu);
  }
  __syncthreads();

  // column t: elite_Q[e, t] for e = 0..k-1 (rank order); mean then population std (tf.math.reduce_std).  The elite rows are
  // regenerated from the counter-based noise: one Philox block (4 consecutive steps of one elite) per thread into shared
  // memory, then every column sums its k entries in rank order (same arithmetic as a per-column loop, 64x fewer Philox calls)
  constexpr int kQCap = 8192;
  __shared__ float sh_q[kQCap];
  float new_mu = 0.0f, new_sd = 0.0f, first_q = 0.0f;
  const int t = threadIdx.x;
  const int H4 = (a.H + 3) >> 2, Hs = H4 * 4;
  if (a.k * Hs <= kQCap) {
    for (int item = threadIdx.x; item < a.k * H4; item += blockDim.x) {
      const int e = item / H4, blk = item - e * H4;
      float zz[4];
      noise4(a.noise, sh_elite[e], (uint32_t)blk, zz);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tt = blk * 4 + j;
        if (tt < a.H) sh_q[e * Hs + tt] = cem_sample(a.mu[tt], a.sd[tt], zz[j], a.lo, a.hi);
      }
    }
    __syncthreads();
    if (t < a.H) {
      float acc = 0.0f;
      for (int e = 0; e < a.k; ++e) acc += sh_q[e * Hs + t];
      first_q = sh_q[t];
      new_mu = acc / (float)a.k;
      float var = 0.0f;
      for (int e = 0; e < a.k; ++e) {
        const float d = sh_q[e * Hs + t] - new_mu;
        var = fmaf(d, d, var);
      }
      new_sd = sqrtf(var / (float)a.k);
    }
  } else if (t < a.H) {  // k x H too large for the staging buffer: per-column regeneration
    const float mu = a.mu[t], sd = a.sd[t];
    float acc = 0.0f;
    for (int e = 0; e < a.k; ++e) {
      const float q = cem_sample(mu, sd, noise1(a.noise, sh_elite[e], t), a.lo, a.hi);
      if (e == 0) first_q = q;
      acc += q;
    }
    new_mu = acc / (float)a.k;
    float var = 0.0f;
    for (int e = 0; e < a.k; ++e) {
      const float q = cem_sample(mu, sd, noise1(a.noise, sh_elite[e], t), a.lo, a.hi);
      const float d = q - new_mu;
      var = fmaf(d, d, var);
    }
    new_sd = sqrtf(var / (float)a.k);
  }
  __syncthreads();  // every column has read the old mu / sd
  if (t < a.H) {
    if (!a.last) {
      a.mu[t] = new_mu;
      a.sd[t] = new_sd;
    } else {
      // :99-102  stdev = clip(stdev, min, 1e8); shift left, append initial stdev / mid-range mean; u = elite_Q[0,0]
      const float sdc = fminf(fmaxf(new_sd, a.sd_min), 1.0e8f);
      if (t > 0) {
        a.mu[t - 1] = new_mu;
        a.sd[t - 1] = sdc;
      } else {
        if (!a.freeze_prev) a.u_prev[0] = first_q;
        if (a.u_out != nullptr) a.u_out[0] = first_q;
        if (a.host.p != nullptr) { a.host.p[8] = first_q; a.host.p[9] = 0.0f; host_publish(a.host); }
      }
      if (t == a.H - 1) {
        a.mu[t] = (a.lo + a.hi) * 0.5f;
        a.sd[t] = a.sd_init;
      }
    }
  }
}

}  // namespace ctk

// ctk_kernels_mppi.cuh -- K1 fused sample -> rollout -> cost -> block softmin partials, K2 combine / u_nom update.
// Replaces reference optimizer_mppi.py:170-193 (_predict_and_cost) for one tick.
#pragma once
#include "ctk_device.cuh"
#include "ctk_predictor.cuh"

namespace ctk {

// One rollout step shared by the segment loops: interpolated perturbation, clip, stage cost, MPPI correction, predictor.
template <class Pred, int KIND, bool LOG>
__device__ __forceinline__ void mppi_step(const MppiArgs& a, Pred& pred, State& z, float& cosang, float& u_last,
                                          float& jsum, float& corr, float u_nom_t, float y0, float y1, float w0,
                                          float w1, int t, int n, bool active) {
  // Interpolator.py:97-106: delta_u[t] = sum_i y_i W[i,t]  (two non-zero terms)
  const float du = fmaf(y1, w1, __fmul_rn(y0, w0));
  const float u = fminf(fmaxf(__fadd_rn(u_nom_t, du), a.lo), a.hi);  // :186-187
  if (LOG && active) {
    float* p = a.log_traj_soa + (size_t)t * 6 * a.N + n;
    p[0] = z.th; p[a.N] = z.om; p[2 * a.N] = z.c; p[3 * a.N] = z.s; p[4 * (size_t)a.N] = z.x; p[5 * (size_t)a.N] = z.v;
    a.log_Q_soa[(size_t)t * a.N + n] = u;
  }
  jsum += stage_cost<KIND>(z, cosang, u, u_last, a.cost);
  // :154-155  cc_weight * (0.5(1-1/NU) R du^2 + R u du + 0.5 R u^2)
  corr = fmaf(a.cc_weight, fmaf(a.coef_du2, du * du, fmaf(a.R * u, du, a.half_R * (u * u))), corr);
  pred.step(z, u);
  cosang = pred.cos_angle(z);
  u_last = u;
}

template <class Pred, int KIND, bool LOG>
__global__ void __launch_bounds__(128) mppi_rollout_kernel(const MppiArgs a) {
  extern __shared__ float smem[];
  float* sh_unom = smem;                 // [H] shifted nominal
  float* sh_w0 = sh_unom + a.H;          // [period]
  float* sh_w1 = sh_w0 + a.period;       // [period]
  float* sh_red = sh_w1 + a.period;      // [32] reduction scratch
  float* sh_part = sh_red + 32;          // [4][n_ind + 1]

  for (int t = threadIdx.x; t < a.H; t += blockDim.x) sh_unom[t] = a.u_nom[min(t + 1, a.H - 1)];
  for (int j = threadIdx.x; j < a.period; j += blockDim.x) interp_weights(j, a.period, &sh_w0[j], &sh_w1[j]);
  Pred pred(a.ode, a.mlp, sh_part + 4 * (a.n_ind + 1));
  __syncthreads();

  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < a.N;
  const uint32_t ng = (uint32_t)(a.off + (active ? n : 0));

  float S = INFINITY;
  if (active || Pred::kCooperative) {
    State z;
    z.th = a.s0[0]; z.om = a.s0[1]; z.c = a.s0[2]; z.s = a.s0[3]; z.x = a.s0[4]; z.v = a.s0[5];
    float cosang = cosf(z.th);  // spec: E_pot uses cos(angle); for t >= 1 the state's own cosine is that value
    float u_last = a.u_prev[0];
    float jsum = 0.0f, corr = 0.0f;
    const int nlog = active ? n : 0;

    float zz[4];
    noise4(a.noise, ng, 0, zz);
    float y_prev = __fmul_rn(zz[0], a.stdev);  // :173-175  normal * stdev (before interpolation)
    int i = 1;                                  // next inducing point
    int t = 0;
    const int nblk = (a.n_ind + 3) >> 2;
    for (int blk = 0; blk < nblk; ++blk) {
      if (blk > 0) noise4(a.noise, ng, blk, zz);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (blk == 0 && q == 0) continue;
        if (i >= a.n_ind) break;
        const float y_cur = __fmul_rn(zz[q], a.stdev);
        const int t_end = min(i * a.period, a.H);
        for (int j = 0; t < t_end; ++t, ++j)
          mppi_step<Pred, KIND, LOG>(a, pred, z, cosang, u_last, jsum, corr, sh_unom[t], y_prev, y_cur, sh_w0[j],
                                     sh_w1[j], t, nlog, active);
        y_prev = y_cur;
        ++i;
      }
    }
    // tail: t == (n_ind-1)*period == H-1 sits exactly on the last inducing point (weight 1)
    for (; t < a.H; ++t)
      mppi_step<Pred, KIND, LOG>(a, pred, z, cosang, u_last, jsum, corr, sh_unom[t], y_prev, 0.0f, 1.0f, 0.0f, t, nlog,
                                 active);

    if (LOG && active) {
      float* p = a.log_traj_soa + (size_t)a.H * 6 * a.N + n;
      p[0] = z.th; p[a.N] = z.om; p[2 * a.N] = z.c; p[3 * a.N] = z.s; p[4 * (size_t)a.N] = z.x; p[5 * (size_t)a.N] = z.v;
    }
    // Cost_Functions/__init__.py:90-92: mean over H+1 of [stage costs, terminal cost]; optimizer_mppi.py:160
    const float Jt = (jsum + terminal_cost(z, a.cost)) / (float)(a.H + 1);
    S = Jt + corr;
    if (active) a.J[n] = S; else S = INFINITY;
  }

  // ---- block softmin partials (optimizer_mppi.py:163-168, restated per block; exact combine in K2) ----
  const float rho_b = block_min(S, sh_red);
  const float e = (active && S < INFINITY) ? expf((S - rho_b) * a.neg_inv_lbd) : 0.0f;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int P = a.n_ind + 1;
  {
    const float ws = warp_sum(e);
    if (lane == 0) sh_part[w * P] = ws;
  }
  const int nblk = (a.n_ind + 3) >> 2;
  for (int blk = 0; blk < nblk; ++blk) {
    float zz[4];
    noise4(a.noise, ng, blk, zz);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = blk * 4 + q;
      if (i < a.n_ind) {
        const float ws = warp_sum(e * zz[q]);
        if (lane == 0) sh_part[w * P + 1 + i] = ws;
      }
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float* out = a.partials + (size_t)blockIdx.x * (P + 1);
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    float acc = 0.0f;
    for (int ww = 0; ww < nw; ++ww) acc += sh_part[ww * P + c];
    out[1 + c] = acc;
  }
  if (threadIdx.x == 0) out[0] = rho_b;
}

__global__ void __launch_bounds__(256) mppi_combine_kernel(const float* __restrict__ in, int cnt, int n_ind,
                                                           float neg_inv_lbd, float* __restrict__ record_out,
                                                           const MppiFinalize fin) {
  extern __shared__ float smem[];
  float* sh_red = smem;            // [32]
  float* sh_rec = smem + 32;       // [2 + n_ind]
  float* sh_u = sh_rec + 2 + n_ind;  // [H] (finalize)
  const int P = n_ind + 2;

  float mn = INFINITY;
  for (int b = threadIdx.x; b < cnt; b += blockDim.x) mn = fminf(mn, in[(size_t)b * P]);
  const float rho = block_min(mn, sh_red);
  for (int c = 0; c < n_ind + 1; ++c) {
    float acc = 0.0f;
    for (int b = threadIdx.x; b < cnt; b += blockDim.x) {
      const float rb = in[(size_t)b * P];
      const float sc = (rb < INFINITY) ? expf((rb - rho) * neg_inv_lbd) : 0.0f;
      acc = fmaf(sc, in[(size_t)b * P + 1 + c], acc);
    }
    const float tot = block_sum(acc, sh_red);
    if (threadIdx.x == 0) sh_rec[1 + c] = tot;
  }
  if (threadIdx.x == 0) sh_rec[0] = rho;
  __syncthreads();
  if (record_out != nullptr)
    for (int c = threadIdx.x; c < P; c += blockDim.x) record_out[c] = sh_rec[c];
  if (!fin.enable) return;

  for (int t = threadIdx.x; t < fin.H; t += blockDim.x) sh_u[t] = fin.u_nom[min(t + 1, fin.H - 1)];
  __syncthreads();
  const float a = sh_rec[1];
  for (int t = threadIdx.x; t < fin.H; t += blockDim.x) {
    const int seg = t / fin.period, j = t - seg * fin.period;
    float w0, w1;
    interp_weights(j, fin.period, &w0, &w1);
    const float bz0 = sh_rec[2 + seg];
    const float bz1 = (j > 0) ? sh_rec[2 + seg + 1] : 0.0f;
    const float b = (fmaf(bz1, w1, bz0 * w0) * fin.stdev) / a;
    const float un = fminf(fmaxf(sh_u[t] + b, fin.lo), fin.hi);
    fin.u_nom[t] = un;
    if (t == 0) {
      if (!fin.freeze_prev) fin.u_prev[0] = un;
      if (fin.u_out != nullptr) fin.u_out[0] = un;
    }
  }
}

// [R][C] -> [C][R] tiled transpose (logs are produced SoA/coalesced by the rollout kernels and handed out in the
// reference's [N, H+1, ns] / [N, H, nu] layout)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[(size_t)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

}  // namespace ctk

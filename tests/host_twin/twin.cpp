// TEST INFRASTRUCTURE ONLY -- host twin of the per-rollout arithmetic.
// Compiles control_toolkit_b200/csrc/ctk_math.cuh (the SAME __host__ __device__ source the kernels use) for the CPU so
// that the ODE step, the cost functions and the hand-derived RPGD adjoint can be checked against the oracle's torch
// autograd in the CPU test-suite, before any GPU time is spent.  It is never loaded by the product
// (control_toolkit_b200 has no CPU path); tests/test_host_twin.py is its only user.
#include <cstring>

#include "../../control_toolkit_b200/csrc/ctk_derive.h"
#include "../../control_toolkit_b200/csrc/ctk_ode_scaled.cuh"

using namespace ctk;

extern "C" {

// J[n] = trajectory cost of rollout n; traj (optional) [N][H+1][6]
void twin_rollout_cost(const float* s0, const float* Q, int N, int H, const ctk_ode_params* op, const ctk_cost_params* cp,
                       float u_prev, float* J, float* traj) {
  FwdK ode; CostC cost;
  derive_fwd(*op, ode);
  derive_cost(*cp, H, cost);
  for (int n = 0; n < N; ++n) {
    State z{s0[0], s0[1], s0[2], s0[3], s0[4], s0[5]};
    float omc = 1.0f - cosf(z.th), ul = u_prev, sum = 0.f;
    for (int t = 0; t < H; ++t) {
      if (traj) { float* p = traj + ((size_t)n * (H + 1) + t) * 6; p[0]=z.th; p[1]=z.om; p[2]=z.c; p[3]=z.s; p[4]=z.x; p[5]=z.v; }
      const float u = Q[(size_t)n * H + t];
      sum += stage_cost_dyn(cost.kind, z, omc, u, ul, cost);
      ode_step(z, u, ode, omc);
      ul = u;
    }
    if (traj) { float* p = traj + ((size_t)n * (H + 1) + H) * 6; p[0]=z.th; p[1]=z.om; p[2]=z.c; p[3]=z.s; p[4]=z.x; p[5]=z.v; }
    J[n] = (sum + terminal_cost(z, cost)) - cost.shift;
  }
}

// grad[n][t] = dJ_n/dQ[n][t] by the same forward-tape + adjoint sweep as rpgd_grad_kernel
void twin_grad(const float* s0, const float* Q, int N, int H, const ctk_ode_params* op, const ctk_cost_params* cp,
               float u_prev, float* grad) {
  OdeC ode; FwdK fwd; CostC cost;
  derive_ode(*op, ode);
  derive_fwd(*op, fwd);
  derive_cost(*cp, H, cost);
  const float w = cost.inv_Hp1;
  State* tape = new State[H];
  for (int n = 0; n < N; ++n) {
    const float* q = Q + (size_t)n * H;
    State z{s0[0], s0[1], s0[2], s0[3], s0[4], s0[5]};
    float omc_unused;
    for (int t = 0; t < H; ++t) { tape[t] = z; ode_step(z, q[t], fwd, omc_unused); }
    Adj lam{0.f, 0.f, 0.f, 0.f};
    for (int t = H - 1; t >= 0; --t) {
      const State& zt = tape[t];
      float g = ode_step_adjoint(zt, q[t], ode, lam);
      g += stage_cost_adjoint_u(q[t], t > 0 ? q[t - 1] : u_prev, t < H - 1 ? q[t + 1] : 0.f, t < H - 1, cost, w);
      grad[(size_t)n * H + t] = g;
      if (t > 0) {
        if (cost.kind == 0) stage_cost_adjoint_state<0>(zt, zt.c, zt.s, cost, w, lam);
        else stage_cost_adjoint_state<1>(zt, zt.c, zt.s, cost, w, lam);
      }
    }
  }
  delete[] tape;
}

// the same gradient through the COEFFICIENT form of the adjoint (adjoint_coefficients in the forward pass, adjoint_apply in the
// reverse sweep), as rpgd_grad_kernel does on the device
void twin_grad_coef(const float* s0, const float* Q, int N, int H, const ctk_ode_params* op, const ctk_cost_params* cp,
                    float u_prev, float* grad) {
  OdeC ode; FwdK fwd; CostC cost;
  derive_ode(*op, ode);
  derive_fwd(*op, fwd);
  derive_cost(*cp, H, cost);
  const float w = cost.inv_Hp1;
  const float D2 = ode.h * ode.inv_mL_kp1L * ode.neg_J_fric, KV = ode.kp1 * ode.neg_M_fric, KU = ode.kp1 * ode.u_max;
  AdjCoef* tape = new AdjCoef[H];
  for (int n = 0; n < N; ++n) {
    const float* q = Q + (size_t)n * H;
    State z{s0[0], s0[1], s0[2], s0[3], s0[4], s0[5]};
    float omc_unused;
    for (int t = 0; t < H; ++t) {
      tape[t] = cost.kind == 0 ? adjoint_coefficients<0>(z, q[t], ode, cost, w) : adjoint_coefficients<1>(z, q[t], ode, cost, w);
      ode_step(z, q[t], fwd, omc_unused);
    }
    Adj lam{0.f, 0.f, 0.f, 0.f};
    for (int t = H - 1; t >= 0; --t) {
      float g = adjoint_apply(tape[t], ode.h, D2, KV, KU, t > 0, lam);
      g += stage_cost_adjoint_u(q[t], t > 0 ? q[t - 1] : u_prev, t < H - 1 ? q[t + 1] : 0.f, t < H - 1, cost, w);
      grad[(size_t)n * H + t] = g;
    }
  }
  delete[] tape;
}

// K1's scaled-variable arithmetic (ctk_ode_scaled.cuh + derive_ode_hot): S[n] = total MPPI cost of rollout n (trajectory cost +
// control-cost correction, optimizer_mppi.py:154-161) for given clipped controls u and unclipped perturbations du [N][H];
// traj (optional) [N][H+1][6] in the reference's unscaled variables.
void twin_rollout_scaled(const float* s0, const float* U, const float* dU, int N, int H, const ctk_ode_params* op, const ctk_cost_params* cp,
                         float u_prev, float cc_weight, float coef_du2, float R, float half_R, float* S, float* traj) {
  OdeHot k;
  MppiCorr mc{(double)cc_weight, (double)coef_du2, (double)R, (double)half_R, -0.01, 1.0, -1.0, 1.0};
  derive_ode_hot(*op, *cp, H, mc, k);
  for (int n = 0; n < N; ++n) {
    ScaledState r;
    scaled_from_state(s0, k, r);
    float ul = u_prev, acc = (k.k_ccrc * u_prev) * u_prev;
    for (int t = 0; t < H; ++t) {
      if (traj) scaled_to_state(r, k, traj + ((size_t)n * (H + 1) + t) * 6);
      const float u = U[(size_t)n * H + t], du = dU[(size_t)n * H + t];
      acc = cp->kind == 0 ? stage_cost_scaled<0>(acc, r, u, ul, du, k) : stage_cost_scaled<1>(acc, r, u, ul, du, k);
      acc = fmaf(du * du, k.k_du2, acc);  // the kernel uses the per-segment closed form of this sum
      ul = u;
      ode_step_scaled(r, u, k);
    }
    if (traj) scaled_to_state(r, k, traj + ((size_t)n * (H + 1) + H) * 6);
    S[n] = finish_cost_scaled(acc, r, ul, k);
  }
}
}

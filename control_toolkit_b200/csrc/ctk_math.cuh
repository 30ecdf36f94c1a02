// ctk_math.cuh -- per-rollout arithmetic of the hot path: CartPole Euler ODE, stage / terminal cost and their
// adjoints.  The formulas are the build's pinned spec (DESIGN.md section "Spec"; reference call sites
// optimizer_mppi.py:188, optimizer_cem_tf.py:57-58, optimizer_rpgd.py:300-303; cost reduction
// Cost_Functions/__init__.py:74-93).  Everything is __host__ __device__ so that the test-only host twin
// (tests/host_twin) can run the SAME code on the CPU against the oracle's autograd; the product only ever
// instantiates it in kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CTK_HD __host__ __device__ __forceinline__
#else
#define CTK_HD inline
#endif

namespace ctk {

// Device-side constant blocks (derived on the host from ctk_ode_params / ctk_cost_params in float64).
struct OdeC {
  float u_max, kp1_Mm, m, neg_M_fric, neg_J_fric, mg, inv_L, kp1, mL, inv_mL, g, inv_kp1L, h;
  // adjoint-only compounds
  float two_m, two_kp1_mL, kp1_neg_M_fric, g_inv_kp1L, inv_mL_kp1L;
  int isteps;
};

struct CostC {
  int kind;  // 0 default, 1 quadratic_boundary_grad
  float dd_weight, ep_weight, ekp_weight, cc_weight, ccrc_weight, R, MAX_COST;
  float inv_two_thl, thl_095, inv_thl_005, thl_09, thl_01;
  float target_position, target_equilibrium;
  float inv_Hp1;  // 1/(H+1) for the adjoint scaling
};

struct State {
  float th, om, c, s, x, v;  // angle, angleD, angle_cos, angle_sin, position, positionD
};

constexpr float kTwoPiHi = 6.28318548202514648f;      // fp32(2*pi)
constexpr float kTwoPiLo = -1.74845553146951715e-7f;   // 2*pi - fp32(2*pi)
constexpr float kInvTwoPi = 0.159154943091895336f;

// wrap to (-pi, pi]: mathematically atan2(sin(th), cos(th)) (spec), evaluated as th - 2*pi*rint(th/(2*pi)) with a
// two-term 2*pi (error <= 1 ulp of th; atan2f(sinf,cosf) itself is only good to ~3 ulp).
CTK_HD float wrap_angle(float th) {
  float k = rintf(th * kInvTwoPi);
  th = fmaf(-k, kTwoPiHi, th);
  return fmaf(-k, kTwoPiLo, th);
}

CTK_HD void sincos_acc(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
  sincosf(x, s, c);
#else
  *s = sinf(x);
  *c = cosf(x);
#endif
}

CTK_HD float fast_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdividef(a, b);  // <= 2 ulp; b = A is in [0.33, 0.43]
#else
  return a / b;
#endif
}

// One predictor step s_{t+1} = f(s_t, Q).  Explicit Euler with the OLD derivatives, then c = cos, s = sin, wrap.
CTK_HD void ode_step(State& z, float Q, const OdeC& p) {
  const float u = p.u_max * Q;
  for (int i = 0; i < p.isteps; ++i) {
    const float A = fmaf(-p.m, z.c * z.c, p.kp1_Mm);
    const float F = p.neg_M_fric * z.v;
    const float T = p.neg_J_fric * z.om;
    const float inner = (F + u) - (p.mL * (z.om * z.om)) * z.s;
    const float num = fmaf(p.kp1, inner, fmaf(p.mg * z.s, z.c, (T * z.c) * p.inv_L));
    const float vd = fast_div(num, A);
    const float wd = (fmaf(p.g, z.s, fmaf(vd, z.c, T * p.inv_mL))) * p.inv_kp1L;
    z.th = fmaf(z.om, p.h, z.th);
    z.om = fmaf(wd, p.h, z.om);
    z.x = fmaf(z.v, p.h, z.x);
    z.v = fmaf(vd, p.h, z.v);
    z.th = wrap_angle(z.th);
    sincos_acc(z.th, &z.s, &z.c);
  }
}

// Stage cost l(s_t, u_t, u_{t-1}) - MAX_COST.  cos_angle = cos(angle) (the state's own cosine for t >= 1).
template <int KIND>
CTK_HD float stage_cost(const State& z, float cos_angle, float u, float u_prev, const CostC& k) {
  const float d = (z.x - k.target_position) * k.inv_two_thl;
  const float ax = fabsf(z.x);
  const float e = (ax - k.thl_095) * k.inv_thl_005;
  float dd = d * d;
  dd += (ax > k.thl_095) ? (1.0e9f * e) * e : 0.0f;
  const float omc = 1.0f - cos_angle;
  const float ep = (k.target_equilibrium * 0.25f) * (omc * omc);
  const float du = u - u_prev;
  float l = k.dd_weight * dd;
  l = fmaf(k.ep_weight, ep, l);
  if (KIND == 1) l = fmaf(k.ekp_weight, z.om * z.om, l);
  l = fmaf(k.cc_weight * k.R, u * u, l);
  l = fmaf(k.ccrc_weight, du * du, l);
  if (KIND == 1) l += (ax > k.thl_09) ? 1.0e7f : 0.0f;
  return l - k.MAX_COST;
}

CTK_HD float stage_cost_dyn(int kind, const State& z, float cos_angle, float u, float u_prev, const CostC& k) {
  return kind == 0 ? stage_cost<0>(z, cos_angle, u, u_prev, k) : stage_cost<1>(z, cos_angle, u, u_prev, k);
}

CTK_HD float terminal_cost(const State& z, const CostC& k) {
  return (fabsf(z.th) > 0.2f || fabsf(z.x - k.target_position) > k.thl_01) ? 10000.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// Reverse mode (RPGD adjoint; reference optimizer_rpgd.py:306-338 obtains it by autodiff).
// lam = dJ/d(th, om, x, v) of the state AFTER this step; on return lam is w.r.t. the state BEFORE the step
// (dynamics part only) and the return value is dJ/dQ through the dynamics.  (c, s) are functions of the angle.
// ---------------------------------------------------------------------------------------------------------------
struct Adj {
  float th, om, x, v;
};

CTK_HD float ode_step_adjoint(const State& z /*state before the step*/, float Q, const OdeC& p, Adj& lam) {
  // recompute forward intermediates (isteps == 1 only; enforced by the host)
  const float u = p.u_max * Q;
  const float A = fmaf(-p.m, z.c * z.c, p.kp1_Mm);
  const float T = p.neg_J_fric * z.om;
  const float F = p.neg_M_fric * z.v;
  const float inner = (F + u) - (p.mL * (z.om * z.om)) * z.s;
  const float num = fmaf(p.kp1, inner, fmaf(p.mg * z.s, z.c, (T * z.c) * p.inv_L));
  const float invA = 1.0f / A;
  const float vd = num * invA;

  const float gwd = lam.om * p.h;                             // d/d(wd)
  const float gvd = fmaf(gwd, z.c * p.inv_kp1L, lam.v * p.h);  // d/d(vd): direct + through wd
  const float gnum = gvd * invA;
  const float gA = -gnum * vd;
  const float gs = fmaf(gnum, fmaf(p.mg, z.c, -(p.kp1 * p.mL) * (z.om * z.om)), gwd * p.g_inv_kp1L);
  const float gc = fmaf(gnum, fmaf(p.mg, z.s, T * p.inv_L), fmaf(gA, -p.two_m * z.c, gwd * (vd * p.inv_kp1L)));
  const float gT = fmaf(gnum, z.c * p.inv_L, gwd * p.inv_mL_kp1L);
  const float gom_direct = gnum * (-p.two_kp1_mL * (z.om * z.s));
  const float gFu = gnum * p.kp1;  // d/dF == d/du

  Adj o;
  o.th = lam.th + fmaf(gs, z.c, -gc * z.s);
  o.om = lam.om + fmaf(lam.th, p.h, fmaf(gT, p.neg_J_fric, gom_direct));
  o.x = lam.x;
  o.v = lam.v + fmaf(lam.x, p.h, gFu * p.neg_M_fric);
  lam = o;
  return gFu * p.u_max;
}

// d(stage cost)/d(state) accumulated into lam (scaled by w = 1/(H+1)); indicator terms have zero gradient.
template <int KIND>
CTK_HD void stage_cost_adjoint_state(const State& z, float cos_angle, float sin_angle, const CostC& k, float w, Adj& lam) {
  const float d = (z.x - k.target_position) * k.inv_two_thl;
  const float ax = fabsf(z.x);
  const float e = (ax - k.thl_095) * k.inv_thl_005;
  float ddx = 2.0f * d * k.inv_two_thl;
  if (ax > k.thl_095) ddx += copysignf(2.0e9f * e * k.inv_thl_005, z.x);
  lam.x = fmaf(w * k.dd_weight, ddx, lam.x);
  const float omc = 1.0f - cos_angle;
  lam.th = fmaf(w * k.ep_weight, (k.target_equilibrium * 0.25f) * (2.0f * omc * sin_angle), lam.th);
  if (KIND == 1) lam.om = fmaf(w * k.ekp_weight, 2.0f * z.om, lam.om);
}

// d(l_t)/d(u_t) + d(l_{t+1})/d(u_t), scaled by w.  has_next = (t < H-1).
CTK_HD float stage_cost_adjoint_u(float u, float u_prev, float u_next, bool has_next, const CostC& k, float w) {
  float g = 2.0f * (k.cc_weight * k.R) * u + 2.0f * k.ccrc_weight * (u - u_prev);
  if (has_next) g -= 2.0f * k.ccrc_weight * (u_next - u);
  return w * g;
}

}  // namespace ctk

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

// Device-side constant blocks (derived on the host from ctk_ode_params / ctk_cost_params in float64, ctk_derive.h).
struct OdeC {
  float u_max, kp1_Mm, m, neg_M_fric, neg_J_fric, mg, inv_L, kp1, mL, inv_mL, g, inv_kp1L, h;
  float kTl, kTm, hk;  // neg_J_fric/L, neg_J_fric/(m L), h/((k+1) L): merged forward constants
  // adjoint-only compounds
  float two_m, two_kp1_mL, kp1_neg_M_fric, g_inv_kp1L, inv_mL_kp1L;
  int isteps;
};

// Forward-pass constants of the ODE, divided through by the pole mass m (9 instead of 12, and u_max folded in):
//   vd = num / A,  num/m = cF v + cU Q - (kp1L w^2) s + g s c + (cTl w) c,   A/m = K1p - c^2
struct FwdK {
  float K1p, kp1L, cF, cU, g, cTl, kTm, h, hk;
  float k0s, k0c;  // leading sincos_half coefficients (register-resident copies)
  int isteps;
};

struct CostC {
  int kind;  // 0 default, 1 quadratic_boundary_grad
  float dd_weight, ep_weight, ekp_weight, cc_weight, ccrc_weight, R, MAX_COST;
  float inv_two_thl, thl_095, inv_thl_005, thl_09, thl_01;
  float target_position, target_equilibrium;
  float inv_Hp1;  // 1/(H+1): the trajectory cost is a MEAN over H+1 terms (Cost_Functions/__init__.py:92)
  // forward constants, pre-multiplied by 1/(H+1) so the rollout accumulates the mean directly
  float k_dd, k_bar, k_ep, k_ekp, k_cc, k_ccrc, k_border, k_term, shift;
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

// sin, cos and 1-cos of an angle in [-pi, pi] without range reduction, quadrant logic or branches: minimax
// polynomials (max error 4.3e-9 / 2.2e-10 before rounding) for sqrt(2) sin(x), sqrt(2) cos(x) at the HALF angle
// x = th/2, written in the variable x' = th/sqrt(2); then sin th = sh ch, 1 - cos th = sh^2, cos th = 1 - sh^2.
// 15 FP32-pipe instructions, no MUFU.  k0s / k0c: the two leading coefficients, passed in so that callers can keep them
// register-resident (an FFMA takes only one immediate).
constexpr float kSinHalfLead = 1.6282471904105478e-07f;
constexpr float kCosHalfLead = -1.1513114017702719e-08f;
CTK_HD void sincos_half(float th, float* s, float* c, float* omc, float k0s = kSinHalfLead, float k0c = kCosHalfLead) {
#if defined(CTK_ACCURATE_SINCOS) && defined(__CUDA_ARCH__)
  sincosf(th, s, c);
  *omc = 1.0f - *c;
  return;
#endif
  const float x = 0.70710678118654752f * th;
  const float t = x * x;
  float ps = fmaf(t, k0s, -2.4761327949818224e-05f);
  ps = fmaf(t, ps, 0.002083262661471963f);
  ps = fmaf(t, ps, -0.08333329111337662f);
  const float sh = fmaf(x * t, ps, x);  // sqrt(2) sin(th/2)
  float pc = fmaf(t, k0c, 2.1885084606765304e-06f);
  pc = fmaf(t, pc, -0.0002455138601362705f);
  pc = fmaf(t, pc, 0.01473138015717268f);
  pc = fmaf(t, pc, -0.3535533845424652f);
  const float ch = fmaf(t, pc, 1.4142135381698608f);  // sqrt(2) cos(th/2)
  *omc = sh * sh;
  *c = fmaf(-sh, sh, 1.0f);
  *s = sh * ch;
}

CTK_HD float fast_rcp(float a) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));  // 1 MUFU, <= 1 ulp; a = A is in [0.33, 0.43]
  return r;
#else
  return 1.0f / a;
#endif
}

// One Euler sub-step with the OLD derivatives, then wrap, c = cos, s = sin.  omc returns 1 - cos(angle) of the new
// state (the E_pot cost term needs it).  17 + 4 + 17 FP32-pipe instructions + 1 MUFU.
CTK_HD void ode_substep(State& z, float Q, const FwdK& p, float& omc) {
  const float w2 = z.om * z.om;
  const float gs = p.g * z.s;
  float n = p.cF * z.v;
  n = fmaf(p.cU, Q, n);
  n = fmaf(-(p.kp1L * w2), z.s, n);
  n = fmaf(gs, z.c, n);
  n = fmaf(p.cTl * z.om, z.c, n);
  const float Ap = fmaf(-z.c, z.c, p.K1p);
  const float vd = n * fast_rcp(Ap);
  const float X = fmaf(vd, z.c, fmaf(p.kTm, z.om, gs));
  z.th = fmaf(z.om, p.h, z.th);
  z.om = fmaf(X, p.hk, z.om);
  z.x = fmaf(z.v, p.h, z.x);
  z.v = fmaf(vd, p.h, z.v);
  z.th = wrap_angle(z.th);
  sincos_half(z.th, &z.s, &z.c, &omc, p.k0s, p.k0c);
}

// One predictor step s_{t+1} = f(s_t, Q): intermediate_steps Euler sub-steps of h = dt / intermediate_steps.
CTK_HD void ode_step(State& z, float Q, const FwdK& p, float& omc) {
  ode_substep(z, Q, p, omc);
  for (int i = 1; i < p.isteps; ++i) ode_substep(z, Q, p, omc);  // headline config: isteps == 1, loop not taken
}

// acc += (stage cost l(s_t, u_t, u_{t-1})) / (H+1), without the MAX_COST shift (applied once per trajectory:
// CostC::shift).  omc = 1 - cos(angle).
template <int KIND>
CTK_HD void stage_cost_acc(float& acc, const State& z, float omc, float u, float u_prev, const CostC& k) {
  // square first, then one FMA with the (uniform) weight: keeps every FMA at <= 2 distinct register sources, which is
  // what the sm_100 register file can feed at full rate (3 distinct register sources issue at ~58 %, measured)
  const float d = z.x - k.target_position;
  const float ax = fabsf(z.x);
  const float e = fmaxf(ax - k.thl_095, 0.0f);  // == indicator(|x| > 0.95 THL) * (|x| - 0.95 THL)
  const float du = u - u_prev;
  acc = fmaf(d * d, k.k_dd, acc);
  acc = fmaf(e * e, k.k_bar, acc);
  acc = fmaf(omc * omc, k.k_ep, acc);
  if (KIND == 1) acc = fmaf(z.om * z.om, k.k_ekp, acc);
  acc = fmaf(u * u, k.k_cc, acc);
  acc = fmaf(du * du, k.k_ccrc, acc);
  if (KIND == 1) acc += (ax > k.thl_09) ? k.k_border : 0.0f;
}

template <int KIND>
CTK_HD float stage_cost(const State& z, float omc, float u, float u_prev, const CostC& k) {
  float l = 0.0f;
  stage_cost_acc<KIND>(l, z, omc, u, u_prev, k);
  return l;
}

CTK_HD float stage_cost_dyn(int kind, const State& z, float omc, float u, float u_prev, const CostC& k) {
  return kind == 0 ? stage_cost<0>(z, omc, u, u_prev, k) : stage_cost<1>(z, omc, u, u_prev, k);
}

// terminal cost / (H+1)
CTK_HD float terminal_cost(const State& z, const CostC& k) {
  return (fabsf(z.th) > 0.2f || fabsf(z.x - k.target_position) > k.thl_01) ? k.k_term : 0.0f;
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
  const float invA = fast_rcp(A);
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

// ---------------------------------------------------------------------------------------------------------------
// The same adjoint step in COEFFICIENT form.  ode_step_adjoint is linear in lam and everything else in it depends only on
// the pre-step state, so the forward pass can fold the state-dependent part into five coefficients (+ the three
// state-gradient terms of the stage cost).  The reverse sweep is then a 4-deep FMA chain per step instead of ~40 dependent
// instructions -- the RPGD tick is ONE warp walking a serial chain, so chain depth is all that matters.
//   gnum = E1 lam.v + E2 lam.om
//   th' = lam.th + C1 gnum + C2 lam.om (+ cth)      om' = lam.om + h lam.th + D1 gnum + D2 lam.om (+ com)
//   x'  = lam.x (+ cx)                              v'  = lam.v + h lam.x + KV gnum          dJ/dQ = KU gnum
// ---------------------------------------------------------------------------------------------------------------
struct AdjCoef {
  float E1, E2, C1, C2, D1;  // dynamics
  float cx, cth, com;        // d(stage cost)/d(x, angle, angleD) of the pre-step state, scaled by w
};

template <int KIND>
CTK_HD AdjCoef adjoint_coefficients(const State& z, float Q, const OdeC& p, const CostC& k, float w) {
  // forward intermediates exactly as in ode_step_adjoint
  const float u = p.u_max * Q;
  const float A = fmaf(-p.m, z.c * z.c, p.kp1_Mm);
  const float T = p.neg_J_fric * z.om;
  const float F = p.neg_M_fric * z.v;
  const float om2 = z.om * z.om;
  const float inner = (F + u) - (p.mL * om2) * z.s;
  const float num = fmaf(p.kp1, inner, fmaf(p.mg * z.s, z.c, (T * z.c) * p.inv_L));
  const float invA = fast_rcp(A);
  const float vd = num * invA;
  AdjCoef r;
  r.E1 = invA * p.h;
  r.E2 = r.E1 * (z.c * p.inv_kp1L);
  const float Ps = fmaf(p.mg, z.c, -(p.kp1 * p.mL) * om2);
  const float Pc = fmaf(p.two_m * z.c, vd, fmaf(p.mg, z.s, T * p.inv_L));
  r.C1 = fmaf(Ps, z.c, -Pc * z.s);
  r.C2 = p.h * fmaf(p.g_inv_kp1L, z.c, -(vd * p.inv_kp1L) * z.s);
  r.D1 = fmaf(z.c * p.inv_L, p.neg_J_fric, -p.two_kp1_mL * (z.om * z.s));
  // stage_cost_adjoint_state
  const float d = (z.x - k.target_position) * k.inv_two_thl;
  const float ax = fabsf(z.x);
  const float e = (ax - k.thl_095) * k.inv_thl_005;
  float ddx = 2.0f * d * k.inv_two_thl;
  if (ax > k.thl_095) ddx += copysignf(2.0e9f * e * k.inv_thl_005, z.x);
  r.cx = (w * k.dd_weight) * ddx;
  const float omc = 1.0f - z.c;
  r.cth = (w * k.ep_weight) * ((k.target_equilibrium * 0.25f) * (2.0f * omc * z.s));
  r.com = (KIND == 1) ? (w * k.ekp_weight) * (2.0f * z.om) : 0.0f;
  return r;
}

// lam (w.r.t. the state after the step) -> lam w.r.t. the state before it, including that state's stage-cost gradient if
// add_cost; returns dJ/dQ through the dynamics.  D2 = h * inv_mL_kp1L * neg_J_fric, KV = kp1 * neg_M_fric, KU = kp1 * u_max.
CTK_HD float adjoint_apply(const AdjCoef& c, float h, float D2, float KV, float KU, bool add_cost, Adj& lam) {
  const float gnum = fmaf(c.E1, lam.v, c.E2 * lam.om);
  Adj o;
  o.th = fmaf(c.C1, gnum, fmaf(c.C2, lam.om, lam.th));
  o.om = fmaf(c.D1, gnum, fmaf(D2, lam.om, fmaf(h, lam.th, lam.om)));
  o.x = lam.x;
  o.v = fmaf(KV, gnum, fmaf(h, lam.x, lam.v));
  if (add_cost) { o.th += c.cth; o.om += c.com; o.x += c.cx; }
  lam = o;
  return KU * gnum;
}

// d(l_t)/d(u_t) + d(l_{t+1})/d(u_t), scaled by w.  has_next = (t < H-1).
CTK_HD float stage_cost_adjoint_u(float u, float u_prev, float u_next, bool has_next, const CostC& k, float w) {
  float g = 2.0f * (k.cc_weight * k.R) * u + 2.0f * k.ccrc_weight * (u - u_prev);
  if (has_next) g -= 2.0f * k.ccrc_weight * (u_next - u);
  return w * g;
}

}  // namespace ctk

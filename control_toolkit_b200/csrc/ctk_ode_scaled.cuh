// ctk_ode_scaled.cuh -- the per-step arithmetic of K1 (ctk_kernels_mppi_ode.cuh) in the scaled state variables of OdeHot:
// T = angle/sqrt(2), W = beta*angleD, V = (cF/g)*positionD.  __host__ __device__ so that the test-only host twin
// (tests/host_twin) runs the SAME source on the CPU against the oracle spec; the product only instantiates it in kernels.
#pragma once
#include "ctk_args.cuh"
#include "ctk_math.cuh"

namespace ctk {

constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kSqrt2 = 1.41421356237309505f;
// period of T = angle/sqrt(2):  2*pi/sqrt(2) = sqrt(2)*pi, split hi + lo
constexpr float kTPerHi = 4.44288301467895508f;      // fp32(sqrt(2)*pi)
constexpr float kTPerLo = -7.6520588976e-08f;         // sqrt(2)*pi - fp32(sqrt(2)*pi)
constexpr float kInvTPer = 0.22507907903927651f;     // 1/(sqrt(2)*pi)

struct ScaledState {
  float T, W, c, s, x, V, omc;  // omc = 1 - cos(angle)
};

CTK_HD void scaled_from_state(const float* s0, const OdeHot& k, ScaledState& r) {
  r.T = s0[0] * kInvSqrt2; r.W = s0[1] * k.beta; r.c = s0[2]; r.s = s0[3]; r.x = s0[4]; r.V = s0[5] * k.cFg;
  r.omc = 1.0f - cosf(s0[0]);  // spec: E_pot uses cos(angle) of the measured state
}
CTK_HD void scaled_to_state(const ScaledState& r, const OdeHot& k, float* out) {
  out[0] = r.T * kSqrt2; out[1] = r.W * k.inv_beta; out[2] = r.c; out[3] = r.s; out[4] = r.x; out[5] = r.V * k.inv_cFg;
}

// stage cost / (H+1) of (state, u, u_prev, du) in the merged form of K1:
//   dd + barrier + E_pot (+ E_kin, border for quadratic_boundary_grad) + u (kA u + kB u_prev + kC du)
// (the u_prev^2 terms of the control-change cost are telescoped into kA; boundary terms are added by the caller)
template <int KIND>
CTK_HD float stage_cost_scaled(float acc, const ScaledState& r, float u, float ul, float du, const OdeHot& k) {
  const float d = r.x - k.target;
  const float e = fmaxf(fabsf(r.x) - k.thl_095, 0.0f);  // indicator(|x| > 0.95 THL) * (|x| - 0.95 THL)
  acc = fmaf(d * d, k.k_dd, acc);
  acc = fmaf(e * e, k.k_bar, acc);
  acc = fmaf(r.omc * r.omc, k.k_ep, acc);
  if (KIND == 1) {
    acc = fmaf(r.W * r.W, k.k_ekp2, acc);
    acc += (fabsf(r.x) > k.thl_09) ? k.k_border : 0.0f;
  }
  float q = u * k.kA;
  q = fmaf(ul, k.kB, q);
  q = fmaf(du, k.kC, q);
  return fmaf(u, q, acc);
}

// One Euler step with the old derivatives (12 FP32 instructions + 1 MUFU), wrap (3), half-angle sincos (14).
CTK_HD void ode_step_scaled(ScaledState& r, float u, const OdeHot& k) {
  float nn = fmaf(k.cUg, u, r.V);
  const float ws = r.W * r.s;
  nn = fmaf(-r.W, ws, nn);
  const float t3 = fmaf(k.cTl2, r.W, r.s);
  nn = fmaf(t3, r.c, nn);
  const float Ap = fmaf(-r.c, r.c, k.K1p);
  const float vd = nn * fast_rcp(Ap);
  const float X = fmaf(vd, r.c, fmaf(k.kTm2, r.W, r.s));
  float T = fmaf(r.W, k.h_T, r.T);
  r.W = fmaf(X, k.h_W, r.W);
  r.x = fmaf(r.V, k.h_x, r.x);
  r.V = fmaf(vd, k.h_V, r.V);
  // wrap: angle <- atan2(sin, cos)  ==  T - period * rint(T / period)
  const float kk = rintf(T * kInvTPer);
  T = fmaf(-kk, kTPerHi, T);
  T = fmaf(-kk, kTPerLo, T);
  r.T = T;
  // half-angle sincos (ctk_math.cuh sincos_half with x = T)
  const float tt = T * T;
  float ps = fmaf(tt, kSinHalfLead, -2.4761327949818224e-05f);
  ps = fmaf(tt, ps, 0.002083262661471963f);
  ps = fmaf(tt, ps, -0.08333329111337662f);
  const float sh = fmaf(T * tt, ps, T);
  float pc = fmaf(tt, kCosHalfLead, 2.1885084606765304e-06f);
  pc = fmaf(tt, pc, -0.0002455138601362705f);
  pc = fmaf(tt, pc, 0.01473138015717268f);
  pc = fmaf(tt, pc, -0.3535533845424652f);
  const float ch = fmaf(tt, pc, 1.4142135381698608f);
  r.omc = sh * sh;
  r.c = fmaf(-sh, sh, 1.0f);
  r.s = sh * ch;
}

// terminal cost / (H+1) and the per-rollout boundary terms of the telescoped control-change cost
CTK_HD float finish_cost_scaled(float acc, const ScaledState& r, float u_last, const OdeHot& k) {
  const float th = r.T * kSqrt2;
  const float term = (fabsf(th) > 0.2f || fabsf(r.x - k.target) > k.thl_01) ? k.k_term : 0.0f;
  return (fmaf(-k.k_ccrc * u_last, u_last, acc) + term) - k.shift;
}

}  // namespace ctk

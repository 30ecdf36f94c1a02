// ctk_derive.h -- host-side derivation of the device constant blocks (OdeC / CostC) from the C-ABI parameter blocks.
// Reciprocals and compounds are evaluated in float64 from the fp32-rounded inputs and rounded once.
#pragma once
#include "../../include/ctk_b200.h"
#include <cmath>

#include "ctk_args.cuh"
#include "ctk_math.cuh"

namespace ctk {

inline void derive_ode(const ctk_ode_params& p, OdeC& o) {
  o.u_max = p.u_max; o.kp1_Mm = p.kp1_Mm; o.m = p.m; o.neg_M_fric = p.neg_M_fric; o.neg_J_fric = p.neg_J_fric;
  o.mg = p.mg; o.kp1 = p.kp1; o.mL = p.mL; o.g = p.g; o.h = p.h;
  o.inv_L = (float)(1.0 / (double)p.L);
  o.inv_mL = (float)(1.0 / (double)p.mL);
  o.inv_kp1L = (float)(1.0 / (double)p.kp1L);
  o.two_m = 2.0f * p.m;
  o.two_kp1_mL = (float)(2.0 * (double)p.kp1 * (double)p.mL);
  o.kp1_neg_M_fric = (float)((double)p.kp1 * (double)p.neg_M_fric);
  o.g_inv_kp1L = (float)((double)p.g / (double)p.kp1L);
  o.inv_mL_kp1L = (float)(1.0 / ((double)p.mL * (double)p.kp1L));
  o.kTl = (float)((double)p.neg_J_fric / (double)p.L);
  o.kTm = (float)((double)p.neg_J_fric / (double)p.mL);
  o.hk = (float)((double)p.h / (double)p.kp1L);
  o.isteps = p.intermediate_steps < 1 ? 1 : p.intermediate_steps;
}
inline void derive_fwd(const ctk_ode_params& p, FwdK& f) {
  const double m = (double)p.m;
  f.K1p = (float)((double)p.kp1_Mm / m);
  f.kp1L = (float)((double)p.kp1 * (double)p.mL / m);
  f.cF = (float)((double)p.kp1 * (double)p.neg_M_fric / m);
  f.cU = (float)((double)p.kp1 * (double)p.u_max / m);
  f.g = (float)((double)p.mg / m);
  f.cTl = (float)((double)p.neg_J_fric / ((double)p.L * m));
  f.kTm = (float)((double)p.neg_J_fric / (double)p.mL);
  f.h = p.h;
  f.hk = (float)((double)p.h / (double)p.kp1L);
  f.k0s = kSinHalfLead;
  f.k0c = kCosHalfLead;
  f.isteps = p.intermediate_steps < 1 ? 1 : p.intermediate_steps;
}
inline void derive_cost(const ctk_cost_params& p, int H, CostC& c) {
  c.kind = p.kind; c.dd_weight = p.dd_weight; c.ep_weight = p.ep_weight; c.ekp_weight = p.ekp_weight;
  c.cc_weight = p.cc_weight; c.ccrc_weight = p.ccrc_weight; c.R = p.R; c.MAX_COST = p.MAX_COST;
  c.inv_two_thl = (float)(1.0 / (double)p.two_thl);
  c.thl_095 = p.thl_095;
  c.inv_thl_005 = (float)(1.0 / (double)p.thl_005);
  c.thl_09 = p.thl_09; c.thl_01 = p.thl_01;
  c.target_position = p.target_position; c.target_equilibrium = p.target_equilibrium;
  const double w = 1.0 / (double)(H + 1);
  c.inv_Hp1 = (float)w;
  const double i2 = 1.0 / (double)p.two_thl, ib = 1.0 / (double)p.thl_005;
  c.k_dd = (float)((double)p.dd_weight * i2 * i2 * w);
  c.k_bar = (float)((double)p.dd_weight * 1.0e9 * ib * ib * w);
  c.k_ep = (float)((double)p.ep_weight * (double)p.target_equilibrium * 0.25 * w);
  c.k_ekp = (float)((double)p.ekp_weight * w);
  c.k_cc = (float)((double)p.cc_weight * (double)p.R * w);
  c.k_ccrc = (float)((double)p.ccrc_weight * w);
  c.k_border = (float)(1.0e7 * w);
  c.k_term = (float)(1.0e4 * w);
  c.shift = (float)((double)p.MAX_COST * (double)H * w);
}

// Scaled-variable constants of the MPPI/ODE kernel (OdeHot, ctk_kernels_mppi_ode.cuh), evaluated in float64.
//   m-divided forward equations (derive_fwd):  n = cF v + cU Q - kp1L w^2 s + g s c + cTl w c,  A = K1p - c^2, vd = n/A,
//   X = vd c + kTm w + g s;  th += h w, w += hk X, x += h v, v += h vd.
//   With W = beta w (beta^2 = kp1L/g), V = (cF/g) v, T = th/sqrt(2) and everything divided by g:
//   n' = V + cUg Q - W^2 s + c (s + cTl2 W),  vd' = n'/A,  X' = vd' c + kTm2 W + s,
//   T += h_T W,  W += h_W X',  x += h_x V,  V += h_V vd'.
struct MppiCorr { double cc_weight, coef_du2, R, half_R, neg_inv_lbd, stdev, lo, hi; };
inline void derive_ode_hot(const ctk_ode_params& p, const ctk_cost_params& cp, int H, const MppiCorr& mc, OdeHot& k) {
  const double m = (double)p.m;
  const double K1p = (double)p.kp1_Mm / m, kp1L = (double)p.kp1 * (double)p.mL / m, cF = (double)p.kp1 * (double)p.neg_M_fric / m;
  const double cU = (double)p.kp1 * (double)p.u_max / m, g = (double)p.mg / m, cTl = (double)p.neg_J_fric / ((double)p.L * m);
  const double kTm = (double)p.neg_J_fric / (double)p.mL, h = (double)p.h, hk = (double)p.h / (double)p.kp1L;
  const double beta = std::sqrt(kp1L / g);
  k.cUg = (float)(cU / g); k.cTl2 = (float)(cTl / (g * beta)); k.kTm2 = (float)(kTm / (g * beta)); k.K1p = (float)K1p;
  k.h_T = (float)(h / (beta * 1.4142135623730951)); k.h_W = (float)(beta * hk * g); k.h_x = (float)(h * g / cF); k.h_V = (float)(cF * h);
  k.beta = (float)beta; k.inv_beta = (float)(1.0 / beta); k.cFg = (float)(cF / g); k.inv_cFg = (float)(g / cF);
  CostC c;
  derive_cost(cp, H, c);
  const double w = 1.0 / (double)(H + 1);
  const double k_cc = (double)cp.cc_weight * (double)cp.R * w, k_ccrc = (double)cp.ccrc_weight * w;
  k.k_dd = c.k_dd; k.k_bar = c.k_bar; k.k_ep = c.k_ep; k.k_ekp2 = (float)((double)cp.ekp_weight * w / (beta * beta));
  k.kA = (float)(k_cc + mc.cc_weight * mc.half_R + 2.0 * k_ccrc);
  k.kB = (float)(-2.0 * k_ccrc);
  k.kC = (float)(mc.cc_weight * mc.R);
  k.k_du2 = (float)(mc.cc_weight * mc.coef_du2);
  k.k_ccrc = (float)k_ccrc;
  k.target = cp.target_position; k.thl_095 = cp.thl_095; k.thl_09 = cp.thl_09; k.k_border = c.k_border; k.thl_01 = cp.thl_01;
  k.k_term = c.k_term; k.shift = c.shift;
  k.lo = (float)mc.lo; k.hi = (float)mc.hi; k.stdev = (float)mc.stdev; k.neg_inv_lbd = (float)mc.neg_inv_lbd;
}

}  // namespace ctk

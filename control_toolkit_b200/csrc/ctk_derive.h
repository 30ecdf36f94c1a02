// ctk_derive.h -- host-side derivation of the device constant blocks (OdeC / CostC) from the C-ABI parameter blocks.
// Reciprocals and compounds are evaluated in float64 from the fp32-rounded inputs and rounded once.
#pragma once
#include "../../include/ctk_b200.h"
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


}  // namespace ctk

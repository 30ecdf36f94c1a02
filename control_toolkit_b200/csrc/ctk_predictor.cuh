// ctk_predictor.cuh -- device predictors: s_{t+1} = f(s_t, Q_t)  (replaces PredictorWrapper.predict_core,
// reference call sites optimizer_mppi.py:188, optimizer_cem_tf.py:57, optimizer_rpgd.py:300).
//   OdePred     : CartPole Euler ODE, one thread per rollout, state in registers.
//   MlpSimtPred : 6 -> hidden tanh -> hidden tanh -> 5 MLP on the FP32 pipe (one thread per rollout, weights in
//                 shared memory).  Numerical anchor for the tcgen05 engine (ctk_mlp_tc.cuh).
#pragma once
#include "ctk_args.cuh"
#include "ctk_device.cuh"

namespace ctk {

struct OdePred {
  static constexpr bool kCooperative = false;
  static constexpr int kRolloutsPerBlock = 0;  // 0: every thread of the block carries a rollout
#ifndef CTK_ODE_MAX_THREADS
#define CTK_ODE_MAX_THREADS 1024
#endif
  static constexpr int kMaxThreads = CTK_ODE_MAX_THREADS;
  static constexpr int kMinBlocks = 1;
  FwdK p;  // forward constants, register-resident (ctk_device.cuh load_fwd)
  __device__ __forceinline__ OdePred(const DevConsts* kc, const MlpDev&, float*) : p(load_fwd(kc)) {}
  // take the FMA-multiplier constants from kernel parameters (uniform registers) instead of vector registers
  __device__ __forceinline__ void use_uniform(const HotUK& u) { p.cU = u.cU; p.kTm = u.kTm; p.h = u.h; p.hk = u.hk; p.K1p = u.K1p; }
  // omc = 1 - cos(angle) of the new state (E_pot cost term)
  __device__ __forceinline__ void step(State& z, float u, float& omc) const { ode_step(z, u, p, omc); }
  // intermediate_steps == 1 fast path (no sub-step loop in the rollout's inner loop)
  __device__ __forceinline__ void substep(State& z, float u, float& omc) const { ode_substep(z, u, p, omc); }
  __device__ __forceinline__ bool single_substep() const { return p.isteps == 1; }
  static size_t smem_floats(const MlpDev&) { return 0; }
};

// tanh with ~1e-7 absolute error on the FP32 pipe + 2 MUFU ops: tanh(x) = 1 - 2/(exp(2x)+1)
__device__ __forceinline__ float tanh_acc(float x) {
  const float ax = fminf(fabsf(x), 15.0f);
  const float e = __expf(2.0f * ax);
  const float r = 1.0f - __fdividef(2.0f, e + 1.0f);
  // small |x|: the formula above cancels; use the odd series (|x| < 0.06: error < 1e-9)
  const float x2 = x * x;
  const float ser = x * fmaf(x2, fmaf(x2, 0.13333334f, -0.33333334f), 1.0f);
  return (ax < 0.06f) ? ser : copysignf(r, x);
}

struct MlpSimtPred {
  static constexpr bool kCooperative = false;
  static constexpr int kRolloutsPerBlock = 0;
  static constexpr int kMaxThreads = 128;
  static constexpr int kMinBlocks = 1;
  int hid;
  const float *W1, *b1, *W2, *b2, *W3T, *b3;  // shared memory
  __device__ __forceinline__ MlpSimtPred(const DevConsts*, const MlpDev& m, float* sm) {
    // 16-byte align the weight block
    float* base = (float*)(((uintptr_t)sm + 15) & ~(uintptr_t)15);
    hid = m.hidden;
    for (int i = threadIdx.x; i < m.blob_floats; i += blockDim.x) base[i] = m.blob[i];
    W1 = base;
    b1 = W1 + 6 * hid;
    W2 = b1 + hid;
    b2 = W2 + hid * hid;
    W3T = b2 + hid;
    b3 = W3T + 5 * hid;
    // caller issues __syncthreads() after construction
  }
  static size_t smem_floats(const MlpDev& m) { return (size_t)m.blob_floats + 4; }
  __device__ __forceinline__ void substep(State& z, float u, float& omc) const { step(z, u, omc); }
  __device__ __forceinline__ bool single_substep() const { return false; }
  __device__ __forceinline__ void use_uniform(const HotUK&) {}

  // net input [Q, angleD, cos, sin, position, positionD] -> next [angleD, cos, sin, position, positionD];
  // angle = atan2(sin, cos)  (oracle/spec.py MLPPredictor.step)
  __device__ __noinline__ void step(State& z, float u, float& omc) const {
    constexpr int HMAX = 128;
    float h1[HMAX];
    const float x[6] = {u, z.om, z.c, z.s, z.x, z.v};
#pragma unroll
    for (int j = 0; j < HMAX; ++j) {
      if (j < hid) {
        float acc = 0.0f;  // torch: (x @ W1) + b1 -> accumulate the dot product first, then add the bias
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fmaf(x[i], W1[i * hid + j], acc);
        h1[j] = tanh_acc(acc + b1[j]);
      }
    }
    float y[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < hid; j0 += 16) {
      float acc[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0f;
#pragma unroll
      for (int i = 0; i < HMAX; ++i) {
        if (i < hid) {
          const float4* wrow = reinterpret_cast<const float4*>(W2 + i * hid + j0);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 wv = wrow[v];
            acc[4 * v + 0] = fmaf(h1[i], wv.x, acc[4 * v + 0]);
            acc[4 * v + 1] = fmaf(h1[i], wv.y, acc[4 * v + 1]);
            acc[4 * v + 2] = fmaf(h1[i], wv.z, acc[4 * v + 2]);
            acc[4 * v + 3] = fmaf(h1[i], wv.w, acc[4 * v + 3]);
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float h2 = tanh_acc(acc[jj] + b2[j0 + jj]);
#pragma unroll
        for (int k = 0; k < 5; ++k) y[k] = fmaf(h2, W3T[k * hid + j0 + jj], y[k]);
      }
    }
    z.om = y[0] + b3[0];
    z.c = y[1] + b3[1];
    z.s = y[2] + b3[2];
    z.x = y[3] + b3[3];
    z.v = y[4] + b3[4];
    z.th = atan2f(z.s, z.c);
    // the network's (cos, sin) outputs are not normalised: cos(atan2(s, c)) = c / hypot(s, c)
    omc = 1.0f - z.c * rsqrtf(fmaf(z.c, z.c, z.s * z.s));
  }
};

}  // namespace ctk

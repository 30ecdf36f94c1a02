// ffma2_lab.cu -- microbenchmarks of sm_100a packed FP32 (fma.rn.f32x2 -> FFMA2) issue rates vs scalar FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_lab ffma2_lab.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// V: 0 scalar FFMA 3 distinct regs (16 chains)   1 scalar FFMA x = x*ku + kb (uniform kernel params)
//    2 FFMA2 three distinct reg pairs (8 chains)  3 FFMA2 x = x*K + z, K from kernel params (uniform?) 4 FMUL2 reg,reg
//    5 FADD2 reg,reg   6 FFMA2 with x = x*y + K(const)   7 mix: FFMA2(3 reg) + scalar FMNMX (alu pipe) 1:1
template <int V>
__global__ void __launch_bounds__(256) lab(float* out, int iters, const float* __restrict__ seed, float ka, float kb) {
  float x[16], y[16], z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = seed[(threadIdx.x + i) & 63];
    y[i] = seed[(threadIdx.x + 16 + i) & 63] * 1e-3f + 0.999f;
    z[i] = seed[(threadIdx.x + 32 + i) & 63] * 1e-3f;
  }
  unsigned long long X[8], Y[8], Z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { X[i] = pack(x[2 * i], x[2 * i + 1]); Y[i] = pack(y[2 * i], y[2 * i + 1]); Z[i] = pack(z[2 * i], z[2 * i + 1]); }
  const unsigned long long KA = pack(ka, ka), KB = pack(kb, kb);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = x[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (V == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], y[i], z[i]);
      } else if (V == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], ka, kb);
      } else if (V == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], Y[i], Z[i]);
      } else if (V == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], KA, Z[i]);
      } else if (V == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) X[i] = mul2(X[i], Y[i]);
      } else if (V == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) X[i] = add2(X[i], Z[i]);
      } else if (V == 6) {
#pragma unroll
        for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], Y[i], KB);
      } else if (V == 7) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { X[i] = fma2(X[i], Y[i], Z[i]); m[i] = fminf(m[i], z[(i + u) & 15]); }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; unpack(X[i], a, b); s += a + b + m[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
static void run(const char* name, double fma_per_thread_iter, float* d, const float* seed) {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    lab<V><<<blocks, threads>>>(d, iters, seed, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double t = fma_per_thread_iter * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (r > 0 && t > best) best = t;
  }
  printf("%-46s %8.2f T fp32-op/s  (x2 = %.1f TFLOP/s if FMA)\n", name, best, 2 * best);
}

int main() {
  float *d, *seed; cudaMalloc(&d, 4 * 148 * 8 * 256); cudaMalloc(&seed, 256);
  float hs[64]; for (int i = 0; i < 64; ++i) hs[i] = 0.5f + 0.001f * i;
  cudaMemcpy(seed, hs, 256, cudaMemcpyHostToDevice);
  run<0>("scalar FFMA 3 distinct regs", 8 * 16, d, seed);
  run<1>("scalar FFMA reg*uniform+uniform", 8 * 16, d, seed);
  run<2>("FFMA2 3 distinct reg pairs", 8 * 16, d, seed);
  run<3>("FFMA2 x*K(param)+z", 8 * 16, d, seed);
  run<4>("FMUL2 reg,reg", 8 * 16, d, seed);
  run<5>("FADD2 reg,reg", 8 * 16, d, seed);
  run<6>("FFMA2 x*y+K(param)", 8 * 16, d, seed);
  run<7>("FFMA2 3reg + FMNMX 1:1 (fp32 ops of FFMA2 only)", 8 * 16, d, seed);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

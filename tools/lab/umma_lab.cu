// umma_lab.cu -- standalone check of the hand-written tcgen05 path used by the MLP predictor engine:
// D[128x128] (fp32, TMEM) = sum over 3-term bf16 splits of A[128x128] . B[128x128]^T, operands in shared memory in the
// canonical no-swizzle K-major core-matrix layout, one thread issuing tcgen05.mma, completion through tcgen05.commit ->
// mbarrier, read-back with tcgen05.ld.  Prints the max error against a float64 CPU product for both LBO/SBO conventions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_lab umma_lab.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// variant 0: LBO = K-direction core-matrix stride, SBO = M/N-direction 8-row-group stride; variant 1: swapped
__global__ void __launch_bounds__(512) umma_test(const __nv_bfloat16* __restrict__ A3, const __nv_bfloat16* __restrict__ B3, float* __restrict__ D,
                                                 int variant, int nterms, long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 3 x 32768
  uint8_t* sB = smem + 3 * 32768;     // 3 x 32768
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands arrive already in the canonical layout: plain copy
  for (int i = tid; i < 3 * 32768 / 16; i += blockDim.x) {
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A3)[i];
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B3)[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_sh)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_sh;
  // instruction descriptor: c=F32 (1<<4), a=BF16 (1<<7), b=BF16 (1<<10), K-major both, N=128 (16<<17), M=128 (8<<24)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  long long t_start = clock64();
  for (int rep = 0; rep < reps; ++rep) {
  if (tid == 0) {
    const uint32_t kstride = 2048, mstride = 128;  // layout: (k/8)*2048 + row*16 + (k%8)*2
    const uint32_t lbo = variant == 0 ? kstride : mstride, sbo = variant == 0 ? mstride : kstride;
    // small terms first: (a3 b1) (a2 b2) (a1 b3) (a2 b1) (a1 b2) (a1 b1)
    const int ta[6] = {2, 1, 0, 1, 0, 0}, tb[6] = {0, 1, 2, 0, 1, 0};
    uint32_t acc = 0;
    for (int term = 6 - nterms; term < 6; ++term) {
      const uint32_t abase = smem_u32(sA + ta[term] * 32768), bbase = smem_u32(sB + tb[term] * 32768);
      for (int ks = 0; ks < 8; ++ks) {
        umma_f16(tmem, make_desc(abase + ks * 2 * kstride, lbo, sbo), make_desc(bbase + ks * 2 * kstride, lbo, sbo), idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"((uint32_t)(rep & 1)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  }
  if (tid == 0 && cycles && blockIdx.x == 0) *cycles = (clock64() - t_start) / reps;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // thread (row) r reads its 128 columns, 32 at a time
  if (tid < 128)
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0) for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

static uint16_t bf16_rn(float f) { __nv_bfloat16 b = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &b, 2); return u; }
static float bf16_f(uint16_t u) { uint32_t x = (uint32_t)u << 16; float f; memcpy(&f, &x, 4); return f; }

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1, block = argc > 2 ? atoi(argv[2]) : 128;
  const int M = 128, N = 128, K = 128;
  std::vector<float> A(M * K), B(N * K);  // B[n][k]
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.09f;
  std::vector<uint16_t> A3(3 * M * K), B3(3 * N * K);
  auto pack = [&](const std::vector<float>& X, std::vector<uint16_t>& X3) {
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 128; ++k) {
        float x = X[r * 128 + k];
        uint16_t t1 = bf16_rn(x); float r1 = x - bf16_f(t1);
        uint16_t t2 = bf16_rn(r1); float r2 = r1 - bf16_f(t2);
        uint16_t t3 = bf16_rn(r2);
        size_t off = (size_t)(k / 8) * 1024 + (size_t)r * 8 + (k % 8);  // in bf16 elements: (k/8)*2048 B + r*16 B + (k%8)*2 B
        X3[0 * 16384 + off] = t1; X3[1 * 16384 + off] = t2; X3[2 * 16384 + off] = t3;
      }
  };
  pack(A, A3); pack(B, B3);
  std::vector<double> ref(M * N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, A3.size() * 2); cudaMalloc(&dB, B3.size() * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, A3.data(), A3.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B3.data(), B3.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 6 * 32768 + 1024;
  cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> D(M * N);
  long long* dC; cudaMalloc(&dC, 8); long long hc = 0;
  for (int variant = 0; variant < 1; ++variant)
    for (int nterms : {1, 3, 6}) {
      cudaMemset(dD, 0, M * N * 4);
      umma_test<<<grid, block, smem>>>(dA, dB, dD, variant, nterms, dC, 200);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d nterms %d: CUDA error %s\n", variant, nterms, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, dC, 8, cudaMemcpyDeviceToHost);
      double mx = 0, mref = 0;
      for (int i = 0; i < M * N; ++i) { mx = fmax(mx, fabs(D[i] - ref[i])); mref = fmax(mref, fabs(ref[i])); }
      printf("[%lld cycles per %d-MMA series] variant %d (LBO=%s) terms %d: max abs err %.3e (max |ref| %.3f)  D[0]=%.6f ref %.6f  D[129]=%.6f ref %.6f\n", hc, nterms * 8, variant,
             variant == 0 ? "K-stride" : "MN-stride", nterms, mx, mref, D[0], ref[0], D[129], ref[129]);
    }
  return 0;
}

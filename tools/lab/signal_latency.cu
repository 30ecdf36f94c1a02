// signal_latency.cu -- measured floor of the serial section of a sharded MPPI tick ("why 8 GPUs stop below 7x").
// u_nom(t+1) depends on EVERY rollout cost of tick t (softmin over the whole population, reference optimizer_mppi.py:163-168,190), so
// between the last rollout of tick t and the first rollout of tick t+1 there is an unavoidable chain of signals:
//   block record -> [L2] -> the GPU's finisher -> [NVLink] -> every peer -> [L2] -> every block of the next tick.
// This tool measures each hop with the same mechanism the kernels use (one 8-byte store carrying value | tag, polled with volatile
// loads): (1) block-to-block ping-pong through L2 on one GPU, (2) a fan-in / fan-out round over G blocks (all publish -> block 0
// polls all -> publishes -> all poll), (3) GPU-to-GPU ping-pong over NVLink peer stores, (4) an all-to-all round over the visible GPUs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/lab/signal_latency.cu -o tools/lab/signal_latency && tools/lab/signal_latency
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void st_tag(unsigned long long* p, unsigned int v, unsigned int tag) {
  const unsigned long long x = ((unsigned long long)tag << 32) | v;
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory");
}
__device__ __forceinline__ void wait_tag(const unsigned long long* p, unsigned int tag) {
  unsigned long long v;
  do { asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); } while ((unsigned int)(v >> 32) != tag);
}

// (1) two blocks of one GPU
__global__ void pingpong_local(unsigned long long* a, unsigned long long* b, int iters, unsigned long long* out) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    const unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) { st_tag(a, i, i); wait_tag(b, i); }
    out[0] = gtime() - t0;
  } else if (blockIdx.x == gridDim.x - 1) {
    for (int i = 1; i <= iters; ++i) { wait_tag(a, i); st_tag(b, i, i); }
  }
}
// (2) fan-in / fan-out round: every block publishes one slot, block 0 polls all of them (one per thread), publishes, all poll
__global__ void round_local(unsigned long long* rec, unsigned long long* bc, int iters, unsigned long long* out) {
  const int G = gridDim.x;
  unsigned long long t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) t0 = gtime();
  for (int i = 1; i <= iters; ++i) {
    if (threadIdx.x == 0) st_tag(rec + blockIdx.x, blockIdx.x, i);
    if (blockIdx.x == 0) {
      for (int b = threadIdx.x; b < G; b += blockDim.x) wait_tag(rec + b, i);
      __syncthreads();
      if (threadIdx.x == 0) st_tag(bc, i, i);
    }
    if (threadIdx.x == 0) wait_tag(bc, i);
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gtime() - t0;
}
// (3) two GPUs: `mine` is local memory the peer stores into, `theirs` is the peer's memory
__global__ void pingpong_peer(unsigned long long* mine, unsigned long long* theirs, int iters, int first, unsigned long long* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long t0 = gtime();
  for (int i = 1; i <= iters; ++i) {
    if (first) { st_tag(theirs, i, i); wait_tag(mine, i); }
    else { wait_tag(mine, i); st_tag(theirs, i, i); }
  }
  out[0] = gtime() - t0;
}
// (4) all-to-all round over W GPUs: store one slot into every GPU's mailbox, poll the W slots of the own mailbox
struct Peers { unsigned long long* box[8]; };
__global__ void round_peer(Peers p, int rank, int W, int iters, unsigned long long* out) {
  if (blockIdx.x != 0) return;
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) t0 = gtime();
  for (int i = 1; i <= iters; ++i) {
    if (threadIdx.x < W) st_tag(p.box[threadIdx.x] + rank, rank, i);
    if (threadIdx.x < W) wait_tag(p.box[rank] + threadIdx.x, i);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = gtime() - t0;
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const int iters = 2000;
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  unsigned long long *slots, *out;
  CK(cudaMalloc(&slots, 4096 * 8)); CK(cudaMemset(slots, 0, 4096 * 8));
  CK(cudaMallocManaged(&out, 64));
  // (1)
  for (int far = 0; far < 2; ++far) {
    CK(cudaMemset(slots, 0, 4096 * 8));
    pingpong_local<<<far ? prop.multiProcessorCount : 2, 32>>>(slots, slots + 16, iters, out);
    CK(cudaDeviceSynchronize());
    printf("(1) block-to-block through L2, %s: round trip %.0f ns, one way %.0f ns\n", far ? "first and last block of a full grid" : "two blocks", (double)out[0] / iters, (double)out[0] / iters / 2);
  }
  // (2)
  for (int G : {8, 37, 74, 148}) {
    if (G > prop.multiProcessorCount) continue;
    CK(cudaMemset(slots, 0, 4096 * 8));
    round_local<<<G, 256>>>(slots, slots + 2048, iters, out);
    CK(cudaDeviceSynchronize());
    printf("(2) fan-in / fan-out round over %3d blocks (publish -> block 0 polls all -> publishes -> all poll): %.0f ns per round\n", G, (double)out[0] / iters);
  }
  if (ndev < 2) { printf("(3), (4): need >= 2 GPUs\n"); return 0; }
  // peer setup
  const int W = ndev > 8 ? 8 : ndev;
  std::vector<unsigned long long*> box(W), res(W);
  for (int d = 0; d < W; ++d) {
    CK(cudaSetDevice(d));
    for (int e = 0; e < W; ++e) if (e != d) { int ok = 0; CK(cudaDeviceCanAccessPeer(&ok, d, e)); if (ok) cudaDeviceEnablePeerAccess(e, 0); }
    cudaGetLastError();
    CK(cudaMalloc(&box[d], 4096)); CK(cudaMemset(box[d], 0, 4096));
    CK(cudaMallocManaged(&res[d], 64));
  }
  // (3)
  for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); pingpong_peer<<<1, 32>>>(box[d] + 64, box[1 - d] + 64, iters, d == 0, res[d]); }
  for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
  printf("(3) GPU-to-GPU over NVLink peer stores: round trip %.0f ns, one way %.0f ns\n", (double)res[0][0] / iters, (double)res[0][0] / iters / 2);
  // (4)
  for (int Wn = 2; Wn <= W; Wn *= 2) {
    Peers p{};
    for (int d = 0; d < Wn; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(box[d], 0, 4096)); CK(cudaDeviceSynchronize()); p.box[d] = box[d]; }
    for (int d = 0; d < Wn; ++d) { CK(cudaSetDevice(d)); round_peer<<<1, 32>>>(p, d, Wn, iters, res[d]); }
    double worst = 0;
    for (int d = 0; d < Wn; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); if ((double)res[d][0] > worst) worst = (double)res[d][0]; }
    printf("(4) all-to-all round over %d GPUs (one tagged slot to every peer, poll the own mailbox): %.0f ns per round\n", Wn, worst / iters);
  }
  return 0;
}

// ctk_topk.cuh -- K4: bitonic top-k on 64-bit (ordered cost, index) keys.
// Replaces tf.argsort(traj_cost)[:k] (reference optimizer_cem_tf.py:73-74, optimizer_rpgd.py:345-346): ascending,
// ties -> lower index first (tf.argsort is top_k(-x)).  One key per thread held in a register; compare-exchange
// partners closer than a warp are reached with __shfl_xor (warp-level bitonic), farther ones through shared memory.
#pragma once
#include "ctk_device.cuh"

namespace ctk {


// sort the 1024 keys of a block ascending; on return thread i holds the i-th smallest.  n_sort (power of two): only the first
// n_sort threads hold real keys (the rest KEY_MAX), so the network can stop at subsequences of that size.
__device__ __forceinline__ uint64_t block_bitonic_sort(uint64_t key, uint64_t* sh /*[1024]*/, int n_sort = TOPK_THREADS) {
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int size = 2; size <= n_sort; size <<= 1) {
    const bool desc = (tid & size) != 0;  // direction of this thread's bitonic subsequence
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      uint64_t other;
      if (stride >= 32) {
        __syncthreads();
        sh[tid] = key;
        __syncthreads();
        other = sh[tid ^ stride];
      } else {
        other = __shfl_xor_sync(0xffffffffu, key, stride);
      }
      const bool lower = (tid & stride) == 0;           // this thread keeps the smaller (asc) of the pair
      const bool take_min = (lower != desc);
      const uint64_t mn = key < other ? key : other, mx = key < other ? other : key;
      key = take_min ? mn : mx;
    }
  }
  return key;
}

}  // namespace ctk

// Grid-wide exclusive scan of one flag per input in a single launch (block scan + last-block-done carry).
#pragma once
#include "ctx.h"

namespace sps {

// Exclusive scan of one 0/1 flag per input over the whole grid: block scan (rank[i] = rank
// inside the block), per-block sums, and the last block to finish turns the sums into exclusive
// block offsets and publishes the total.  Global rank of i = rank[i] + block_sums[i / kScanBlock].
// Must be called by every thread of every block with blockIdx.x < nb.
__device__ __forceinline__ int scan_flags(int flag, int i, int n, int nb, int32_t* __restrict__ rank, int32_t* block_sums,
                                 uint32_t* ticket, int32_t* count_out) {
  __shared__ int warp_sums[kScanBlock / 32];
  __shared__ bool is_last;
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int incl = flag;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += v;
    }
    if (lane < kScanBlock / 32) warp_sums[lane] = wi - w;  // exclusive
  }
  __syncthreads();
  const int excl = incl - flag + warp_sums[wid];
  if (i < n) rank[i] = excl;
  if (tid == kScanBlock - 1) {
    block_sums[blockIdx.x] = excl + flag;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == (uint32_t)(nb - 1));
  }
  __syncthreads();
  if (!is_last) return excl;
  __threadfence();
  // serial-over-chunks exclusive scan of block_sums[0..nb) by this block
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += kScanBlock) {
    const int j = base + tid;
    const int v = (j < nb) ? ((volatile int32_t*)block_sums)[j] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += u;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
      int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += u;
      }
      if (lane < kScanBlock / 32) warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const int ex = inc - v + warp_sums[wid] + carry;
    if (j < nb) block_sums[j] = ex;
    __syncthreads();
    if (tid == kScanBlock - 1) carry = ex + v;
    __syncthreads();
  }
  if (tid == 0) {
    *count_out = carry;
    *ticket = 0;  // ready for the next scan on this stream
  }
  return excl;
}


}  // namespace sps

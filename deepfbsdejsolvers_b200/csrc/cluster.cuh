// Per-path reductions of the jump-scheme kernels: over the G threads of a path inside a CTA, and over the CTAs of a
// thread-block cluster that share one path (small batches).  Shared by pricing_kernels.cu and jump_tc_kernels.cu.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace fbsdej {

__device__ __forceinline__ float group_allsum(float v, int G, float* red) {
  if (G <= 32) return group_sum_shfl(v, G);
  return block_sum(v, red);   // G == kThreads: one path per CTA
}

// two sums at once (one barrier pair for G == kThreads); red[0 .. 7]
__device__ __forceinline__ void group_allsum2(float& u, float& v, int G, float* red) {
  if (G <= 32) { u = group_sum_shfl(u, G); v = group_sum_shfl(v, G); return; }
  u = warp_sum(u); v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = u; red[4 + (threadIdx.x >> 5)] = v; }
  __syncthreads();
  u = (red[0] + red[1]) + (red[2] + red[3]);
  v = (red[4] + red[5]) + (red[6] + red[7]);
}

// Small batches (the reference's B = 10): a thread-block CLUSTER of C CTAs shares one path and splits its compensator
// samples; the per-step partial sums are exchanged through distributed shared memory.  v[] is CTA-uniform on entry
// (after block_sum); every CTA of the cluster leaves with the same sum, added in rank order.  Two slots alternate so
// that one cluster barrier per call suffices (a slot is rewritten only after the next call's barrier).
constexpr int kRedFloats = 8 + 2 * 16;
template <int NV>
__device__ __forceinline__ void cluster_allsum(float (&v)[NV], float* red, int& parity, int C) {
  static_assert(NV <= 16, "slot width");
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  float* slot = red + 8 + 16 * parity;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) slot[k] = v[k];
  }
  cl.sync();
  float s[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) s[k] = 0.0f;
  for (int r = 0; r < C; ++r) {
    const float* __restrict__ rs = cl.map_shared_rank(slot, r);
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k] += rs[k];
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s[k];
  parity ^= 1;
}
__device__ __forceinline__ unsigned cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }


}  // namespace fbsdej

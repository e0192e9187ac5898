// Small utility kernels: partial reduction, Keras-form Adam, generic row-wise net evaluation, layout transposes.
#pragma once
#include "common.cuh"

namespace fbsdej {

// Data-parallel exchange of the step's [loss | gradient] vector, fused into the finishing kernel (util_kernels.cu): every rank
// owns one buffer  data [2 slots][world][nstride] floats | flags [world][nblk] u32 | step counter u32  that its peers can
// write (same process: plain pointers; other processes: CUDA IPC over NVLink).  peers / rank / world live on the host side
// in DpExchange, the kernel sees device tables of the W data and flag pointers.
struct XchgArgs {
  float* const* peer_data;       // device array [world]: data block of every rank (own included)
  uint32_t* const* peer_flags;   // device array [world]: flag block of every rank
  uint32_t* xctr;                // this rank's step stamp (monotonic); xctr[1] = error word: stamp of the first step that timed out
  int rank, world, nstride, nblk;
  unsigned long long timeout_ns; // how long a block waits for a peer's flag before it voids the step
};
__host__ __device__ inline int xchg_nblk(int P) { return (kHeader + P + 31) / 32; }
__host__ __device__ inline int xchg_nstride(int P) { return xchg_nblk(P) * 32; }
inline size_t xchg_bytes(int P, int world) {
  return sizeof(float) * 2 * (size_t)world * xchg_nstride(P) + sizeof(uint32_t) * ((size_t)world * xchg_nblk(P) + 4);
}

int launch_reduce_partials(const float* lpart, int nparts_l, const float* gpart, int nparts_g, int P, float* out,
                           bool with_grad, cudaStream_t st);
int launch_reduce_adam(const float* lpart, int nparts_l, const float* gpart, int nparts_g, int P, float* out, float* theta,
                       float* m, float* v, const float* mask, float lr, float b1, float b2, float eps, int* t_dev,
                       uint32_t* iter_dev, float* loss_dst, uint32_t* step_ctr, unsigned int* done_ctr, cudaStream_t st,
                       const XchgArgs* x = nullptr);
int launch_adam(float* theta, float* m, float* v, const float* grad, const float* mask, int n, float lr, float b1,
                float b2, float eps, int* t_dev, cudaStream_t st);
int launch_bump_u32(uint32_t* p, cudaStream_t st);
int launch_copy_loss(const float* out, float* dst, uint32_t* ctr, cudaStream_t st);
int launch_net_forward(const float* theta_net, int nin, int H, int L, int nout, int act, const float* x, int rows,
                       float* y, cudaStream_t st);
int launch_scatter(float* dst, size_t n, const uint32_t* idx, const float* val, int nnz, cudaStream_t st);
int launch_transpose(const float* src, float* dst, int N, int B, int d, bool to_ndb, cudaStream_t st);

}  // namespace fbsdej

// Small utility kernels: partial reduction, Keras-form Adam, generic row-wise net evaluation, layout transposes.
#pragma once
#include "common.cuh"

namespace fbsdej {

int launch_reduce_partials(const float* lpart, int nparts_l, const float* gpart, int nparts_g, int P, float* out,
                           bool with_grad, cudaStream_t st);
int launch_reduce_adam(const float* lpart, int nparts_l, const float* gpart, int nparts_g, int P, float* out, float* theta,
                       float* m, float* v, const float* mask, float lr, float b1, float b2, float eps, int* t_dev,
                       uint32_t* iter_dev, float* loss_dst, uint32_t* step_ctr, unsigned int* done_ctr, cudaStream_t st);
int launch_adam(float* theta, float* m, float* v, const float* grad, const float* mask, int n, float lr, float b1,
                float b2, float eps, int* t_dev, cudaStream_t st);
int launch_bump_u32(uint32_t* p, cudaStream_t st);
int launch_copy_loss(const float* out, float* dst, uint32_t* ctr, cudaStream_t st);
int launch_net_forward(const float* theta_net, int nin, int H, int L, int nout, int act, const float* x, int rows,
                       float* y, cudaStream_t st);
int launch_transpose(const float* src, float* dst, int N, int B, int d, bool to_ndb, cudaStream_t st);

}  // namespace fbsdej

// Argument blocks of the increment-simulation kernels (sim_kernels.cu).
#pragma once
#include "common.cuh"

namespace fbsdej {

struct SimCommon {};

struct SimMertonArgs {
  int B, N, D;                 // B = paths (or M for the compensator samples)
  uint32_t seed_lo, seed_hi, iteration, path_offset, stream;
  const uint32_t* iter_ptr;    // device iteration counter (CUDA-graph replay); NULL -> `iteration`
  float sqdt, muJ, sigJ;
  const uint32_t* pois_thr;    // device table, npois <= 64 entries
  int npois;
  float* dW;                   // [N][D][B] or NULL
  float* J;                    // [N][D][B]
};

struct SimVGArgs {
  int B, N;
  uint32_t seed_lo, seed_hi, iteration, path_offset, stream;
  const uint32_t* iter_ptr;
  float shape, scale, theta, sigJ;
  float* J;                    // [N][B]
};

struct SimMFGArgs {
  int B, N;
  uint32_t seed_lo, seed_hi, iteration, path_offset, stream;
  const uint32_t* iter_ptr;
  float sqdt, dt, q0, alpha, beta, jumpFactor, coeffOU, sig0;
  int stochastic;
  const float* qaver;          // device [N+1]
  float* dW0; float* dW; float* dN;   // [N][B]
};

int launch_sim_merton(const SimMertonArgs& a, cudaStream_t st);
int launch_sim_vg(const SimVGArgs& a, cudaStream_t st);
int launch_sim_mfg(const SimMFGArgs& a, cudaStream_t st);
int launch_compact_jmc(const float* src, float* dst, int* nnz, int* n0, int N, int D, int M, int dedup, cudaStream_t st);

}  // namespace fbsdej

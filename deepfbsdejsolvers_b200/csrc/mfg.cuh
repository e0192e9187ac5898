// Argument block of the fused MFG kernels (mfg_kernels.cu).
#pragma once
#include "pricing.cuh"

namespace fbsdej {

struct MFGArgs {
  int B, N, scheme;          // SCH_*
  int has_y, has_z;          // nets output Y first (MultiStep/SumLocal/Reg); Z0/Gam(/Z) present (not Reg)
  int stochastic;            // jumpModel == 'stochastic'
  int mma_mode;              // 0 = fp32 FFMA kernels (mfg_kernels.cu), 1 = tcgen05 kernels (mfg_tc_kernels.cu)
  float inv_B, w_hat, w_ind;
  float dt, q0, R0, S0;
  float alpha, beta, jumpFactor, coeffOU, A, K, pi, p0, p1, f0, f1, thetaR, C, h1, h2, sig0, sig, alphaTarget, coeffEqui;
  const float* qaver;        // [N+1]
  const float* meanhq;       // [N+1]  deterministic mean of hQ (MFGModel.py:67-68), host float64 -> fp32
  NetRt netA, netB;
  int y0_off, P;
  const float* theta;
  const float* dW0; const float* dW; const float* dN;   // [N][B]
  float* traj;               // [N+1][5][B]  hQ, Q, R, hS, S
  float* sch;                // [N][2][B]    MultiStep: (e_h, e) ; SumLocal: (rho_h, rho)
  float* fin;                // [2][B]
  float* trajY;              // optional [N+1][2][B]
  float* lpart;              // [grid][4]
  float* gpart;              // [grid][P]
};

int launch_mfg(int HP, const MFGArgs& a, int grid, bool backward, cudaStream_t st);
int launch_mfg_tc(const MFGArgs& a, int grid, bool backward, cudaStream_t st);
int launch_pricing(int model, int D, int HP, const PricingArgs& a, int grid, bool backward, cudaStream_t st);
size_t pricing_smem_bytes(int HP, const PricingArgs& a, bool backward);
size_t mfg_smem_bytes(int HP, const MFGArgs& a, bool backward);
int pricing_blocks_per_sm(int model, int D, int HP, const PricingArgs& a, bool backward);
int mfg_blocks_per_sm(int HP, const MFGArgs& a, bool backward);
int launch_price(int model, int D, const PricingArgs& a, int iStep, const float* X, int n, float* out, cudaStream_t st);
int launch_untile_traj(int D, const float* rec, const float* recN, TileMap map, int B, int N, float* out, cudaStream_t st);
size_t reg_tc_wimg_floats();
size_t reg_tc_wimg_fwd_floats();
int launch_reg_stage_operands(const PricingArgs& a, float* img, cudaStream_t st);

}  // namespace fbsdej

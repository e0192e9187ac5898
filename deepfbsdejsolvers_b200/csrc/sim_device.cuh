// Device-side samplers of the Merton increments, shared by the simulation kernel (sim_kernels.cu) and by the tcgen05
// forward sweep when it draws its own increments (reg_tc_kernels.cu, PricingArgs::rng): identical counters, identical
// arithmetic, hence bit-identical increments whether they are materialised in HBM or consumed in registers.
//
// One (path, asset pair, step) cell = ONE Philox4x32-10 block, counter (global path id, step << 8 | pair, iteration, stream):
// words 0/1 -> the two Brownian normals (Box-Muller), words 2/3 -> the two Poisson counts by table inversion.
#pragma once
#include "common.cuh"

namespace fbsdej {

// Rare cases of a Poisson draw (count >= 2, or a single jump whose uniform falls in the far tail of the normal): exact
// table walk, library inverse CDF, second Philox block for the collapsed sum of count normals (pricingModels.py:60).
static __device__ __noinline__ float jump_size_rare(uint32_t u, uint32_t t0, uint32_t t1, float inv_w1, const uint32_t* __restrict__ thr,
                                             int n, float muJ, float sigJ, uint32_t gid, uint32_t c1, uint32_t iter, uint32_t stream,
                                             uint32_t k0, uint32_t k1, int which) {
  if (u < t1) {
    const float v = ((float)(u - t0) + 0.5f) * inv_w1;
    return fmaf(sigJ, normcdfinvf(fminf(fmaxf(v, 1.0e-9f), 0.99999994f)), muJ);
  }
  int c = 2;
  for (int k = 2; k < n; ++k) {
    if (u < thr[k]) break;
    ++c;
  }
  const uint4 s = Philox::rand4(gid, c1, iter, rare_jump_stream(stream), k0, k1);
  float e0, e1;
  box_muller(s.x, s.y, e0, e1);
  const float dn = (float)c;
  return dn * muJ + sigJ * sqrtf(dn) * (which ? e1 : e0);
}

// MUFU.LG2 without the denormal-input rescaling __log2f adds (three instructions): the arguments here are >= 2^-33 or exactly 0
__device__ __forceinline__ float lg2_raw(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_raw(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Branch-free common case for the two Poisson draws (u0, u1) of one Philox block.  count = 0: J = 0.  count = 1
// (thr[0] <= u < thr[1]): v = (u - thr[0] + 1/2) / (thr[1] - thr[0]) is uniform on (0, 1) with ~28 bits;
// z = Phi^{-1}(v) = sqrt(2) erfinv(2v - 1) by Giles' single-precision polynomial in w = -ln(4 v (1 - v)) (central branch
// w < 5, |error| ~ 3e-7); J = muJ + sigJ z.  The polynomial is evaluated ONCE, on whichever of the two draws has the single
// jump; `rare` flags everything else (a count >= 2, both draws jumping, a far-tail size) for the exact path.
__device__ __forceinline__ void jump_sizes_fast(uint32_t u0, uint32_t u1, uint32_t t0, uint32_t t1, float inv_w1, float muJ,
                                                float sigJ, float& j0, float& j1, uint32_t& rare) {
  const bool one0 = (u0 >= t0) && (u0 < t1), one1 = (u1 >= t0) && (u1 < t1);
  const uint32_t u = one1 ? u1 : u0;
  const float v = ((float)(u - t0) + 0.5f) * inv_w1;
  const float x = fmaf(2.0f, v, -1.0f);
  float w = -0.6931471805599453f * lg2_raw(fmaf(-x, x, 1.0f));
  const bool tail = (one0 || one1) && !(w < 5.0f);
  rare = ((u0 >= t1) || (one0 && (one1 || tail)) ? 1u : 0u) | ((u1 >= t1) || (one1 && (one0 || tail)) ? 2u : 0u);
  w -= 2.5f;
  float p = 2.81022636e-08f;
  p = fmaf(p, w, 3.43273939e-07f);
  p = fmaf(p, w, -3.5233877e-06f);
  p = fmaf(p, w, -4.39150654e-06f);
  p = fmaf(p, w, 0.00021858087f);
  p = fmaf(p, w, -0.00125372503f);
  p = fmaf(p, w, -0.00417768164f);
  p = fmaf(p, w, 0.246640727f);
  p = fmaf(p, w, 1.50140941f);
  const float jump = fmaf(sigJ, 1.4142135623730951f * p * x, muJ);
  j0 = one0 ? jump : 0.0f;
  j1 = one1 ? jump : 0.0f;
}

// two N(0,1) from two 32-bit words: Box-Muller with the MUFU approximations (lg2, sqrt, sin, cos)
__device__ __forceinline__ void box_muller_fast(uint32_t a, uint32_t b, float scale, float& n0, float& n1) {
  const float u = fmaf(__uint2float_rz(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (0, 1)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg2_raw(u)));           // sqrt(-2 ln u)
  r *= scale;
  float sn, cs;
  __sincosf(__uint2float_rz(b) * 1.4629180792671596e-09f, &sn, &cs);                               // 2 pi b / 2^32
  n0 = r * cs;
  n1 = r * sn;
}


// The two (dW, J) pairs of one cell.  sthr: the Poisson thresholds in shared memory (rare path only).
struct MertonCell { float w0, w1, j0, j1; };
__device__ __forceinline__ MertonCell merton_cell(uint32_t gid, uint32_t c1, uint32_t iter, uint32_t stream, uint32_t k0, uint32_t k1,
                                                  uint32_t t0, uint32_t t1, float inv_w1, float sqdt, float muJ, float sigJ,
                                                  const uint32_t* __restrict__ sthr, int npois) {
  MertonCell c;
  const uint4 r = Philox::rand4(gid, c1, iter, stream, k0, k1);
  box_muller_fast(r.x, r.y, sqdt, c.w0, c.w1);
  uint32_t rare;
  jump_sizes_fast(r.z, r.w, t0, t1, inv_w1, muJ, sigJ, c.j0, c.j1, rare);
  if (rare) {                                              // ~0.4 % of the cells: exact path
    if (rare & 1u) c.j0 = jump_size_rare(r.z, t0, t1, inv_w1, sthr, npois, muJ, sigJ, gid, c1, iter, stream, k0, k1, 0);
    if (rare & 2u) c.j1 = jump_size_rare(r.w, t0, t1, inv_w1, sthr, npois, muJ, sigJ, gid, c1, iter, stream, k0, k1, 1);
  }
  return c;
}

}  // namespace fbsdej

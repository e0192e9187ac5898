// Argument block and model policies of the fused pricing-path kernels (Merton jump-diffusion,
// d-asset geometric-basket extension, Variance Gamma).
#pragma once
#include "tile_mlp.cuh"

namespace fbsdej {

// loss-graph families (reference classes: SolversJumpDiff.py / SolversPureJump.py)
enum { SCH_GLOBAL = 0, SCH_MULTISTEP = 1, SCH_SUMLOCAL = 2 };

// Tiles of the tcgen05 compensator-free kernels (reg_tc_kernels.cu).  Tile t holds the paths [tile_base(t), + tile_height(t)):
// 128 rows (four warps) when (t % period) < tall, else 96 rows (three warps); period = the grid size, so a CTA's tiles all
// have its own height.  The sweeps are bound by instruction issue, and 2^16 paths are 13.8 warps per SM: with 128-row tiles
// only, 68 SMs carry 16 warps while 80 carry 12; the mixed heights give every SM 14 (make_tile_map, api.cu).
// Uniform 128-row tiles: period = tall = 1.
struct TileMap { int ntiles, period, tall; };
__host__ __device__ inline int tile_height(const TileMap& m, int t) { return (t % m.period) < m.tall ? 128 : 96; }
__host__ __device__ inline int tile_base(const TileMap& m, int t) {
  const int w = t / m.period, c = t % m.period;
  return w * (m.tall * 128 + (m.period - m.tall) * 96) + (c < m.tall ? c * 128 : m.tall * 128 + (c - m.tall) * 96);
}
inline TileMap make_tile_map(int B, int slots) {
  const long long W = (B + 31) / 32;                          // warps of paths
  const long long waves = (W + 4LL * slots - 1) / (4LL * slots);
  if (W >= 3LL * slots * waves) {                             // every CTA slot gets a tile of 3 or 4 warps in every wave
    const int tall = (int)((W - 3LL * slots * waves + waves - 1) / waves);
    return TileMap{(int)(waves * slots), slots, tall};
  }
  return TileMap{(B + 127) / 128, 1, 1};
}

struct PricingArgs {
  int B, N, G, M;             // local paths, time steps, threads per path, compensator sample count (mean denominator)
  int C;                      // CTAs per path: a thread-block cluster splits the compensator samples (small batches; else 1)
  int scheme;                 // SCH_*
  int one_net;                // jump rows are evaluated by netA (MultiStep1 / SumLocal1)
  int has_jump;               // 0 for the *Reg solvers (no Z / Gam / compensator)
  int use_netA;               // 0 only for VG Global (U network unused, SolversPureJump.py:22-41)
  int has_y, zoff, has_z;     // netA output map: out[0] = Y if has_y; Z[k] = out[zoff + k] if has_z
  int feat_mode;              // Merton two-net: 0 -> J (Global), 1 -> e^J
  int stale_time;             // SumLocal*: time feature of step k >= 1 is k-1
  int mma_mode;               // 0 = fp32 FFMA layers, 1 = tcgen05 (compensator-free solvers only)
  int jump_sep;               // tcgen05 jump rows, one path per thread: separable first layer (jump_tc.cuh; FBSDEJ_NO_JUMP_SEP=1 turns it off)
  float inv_B;                // 1 / GLOBAL batch (data-parallel ranks sum their partial means)
  float dt, r, K, x0, aLin, sig, drift_dt;
  NetRt netA, netB;
  int y0_off, P;
  // rng = 1 (tcgen05 forward, Merton): the sweep draws its own increments (sim_device.cuh: same counters and arithmetic
  // as sim_merton_kernel) instead of reading dW / J
  int rng, npois;
  uint32_t seed_lo, seed_hi, iteration, path_offset;
  const uint32_t* iter_ptr;   // device iteration counter (CUDA-graph replay); NULL -> `iteration`
  const uint32_t* pois_thr;
  float sqdt, muJ, sigJ;
  const float* theta;
  const float* dW;            // [N][D][B]
  const float* J;             // [N][D][B]
  const float* JMC;           // [N][D][Mcap]  non-zero samples first
  const int* jmc_nnz;         // [N]
  const int* jmc_n0;          // [N] multiplicity of the all-zero sample
  int Mcap;
  // Merton series tables (host float64 -> fp32): per (step, n): (c1, c2, sig_n*sqrt(tau), w_n), wK_n
  const float4* tabA;
  const float* tabK;
  const int2* tab_range;      // [N] (nlo, nhi)
  const float* qdisc;         // [N] e^{-q tau_i} (1 when d == 1)
  int limit;
  // optional tabulation of the series in k = log(G e^{-q tau} / K): per step a uniform grid of nodes
  // (sD, dsD/dk, sK, dsK/dk), cubic Hermite between nodes; outside the grid the series is summed term by term
  const float4* atab;         // all steps' nodes, concatenated
  const float4* atab_meta;    // [N] (kmin, 1/h, h, number of intervals)
  const int* atab_off;        // [N] first node of step i
  int use_atab;
  // VG spline tables: per (step, interval) cubic coefficients (c0..c3) around knot k0 + idx*h
  const float4* vg_coef;
  const float* vg_scale;      // [N] e^{-r tau_i} / pi
  int vg_nint;
  float vg_k0, vg_h, vg_inv_h;
  // per path-step stores for the adjoint sweep
  float* trajX;               // [N+1][D][B]
  float* trajE;               // [N][D][B]  exp(drift dt + sig dW + J) (compensator-free solvers; saves the adjoint 2 loads + 1 exp)
  float* aux_s;               // [N][B]  aLin*dt*sign(Ysel - A_i)
  float* aux_dA;              // [N][B]  dA/dX (d=1) or G*dA/dG/d (d>1)
  float* sch1;                // [N][B]  MultiStep: e_k = F_k - g ; SumLocal: rho_i
  float* fin;                 // [B]     Global: Y_N - g ; MultiStep: sum_k e_k
  // tcgen05 (compensator-free) path: tile-major per path-step records, one contiguous block per (tile, step):
  //   rec  [ntiles][N][2D+3][128]   planes X[D], E[D], aLin*dt*sign, dA base, scheme residual  (RecLayout)
  //   recN [ntiles][D+1][128]       planes X_N[D], fin
  float* rec;
  float* recN;
  TileMap tmap;
  // weight operand images of the tcgen05 compensator-free kernels, built from theta by reg_stage_operands_kernel once per pass
  // and bulk-copied (TMA) into shared memory by every CTA: forward (TF32 hi / lo, r-form scales), adjoint (bf16 hi / lo)
  const float* wimg_fwd;
  const float* wimg_bwd;
  float* trajY;               // optional [N+1][B]
  float* trajZ;               // optional [N][D][B]
  float* lpart;               // [grid][4]
  float* gpart;               // [grid][P]
};

// Record layout of the tcgen05 path: the adjoint sweep reads one (tile, step) block with immediate offsets from a
// single base pointer (and one bulk L2 prefetch), the forward sweep writes it with 512-byte coalesced plane rows.
template <int D>
struct RecLayout {
  static constexpr int P_X = 0, P_E = D, P_S = 2 * D, P_DA = 2 * D + 1, P_SCH = 2 * D + 2, NP = 2 * D + 3;
  static constexpr int NPT = D + 1;
  __host__ __device__ static size_t step_floats() { return (size_t)NP * TR; }
  __host__ __device__ static size_t rec_floats(int ntiles, int N) { return (size_t)ntiles * N * NP * TR; }
  __host__ __device__ static size_t recN_floats(int ntiles) { return (size_t)ntiles * NPT * TR; }
};

template <int D_>
struct MertonModel {
  static constexpr int D = D_;
  static constexpr bool kBrownian = true;

  __device__ static __forceinline__ float basket(const float (&X)[D]) {
    if (D == 1) return X[0];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) s += logf(X[k]);
    return expf(s * (1.0f / D));
  }
  // Closed-form price A(i, X) (pricingModels.py:33-49) and the stored derivative base.
  __device__ static __forceinline__ void eval_A(const PricingArgs& a, int i, const float (&X)[D], float& A, float& dAb) {
    const float G = basket(X);
    const float Ge = (D == 1) ? G : G * a.qdisc[i];
    const float k = logf(Ge / a.K);
    float sD = 0.0f, sK = 0.0f;
    bool done = false;
    if (a.use_atab) {
      const float4 m = __ldg(a.atab_meta + i);
      const float u = (k - m.x) * m.y;
      if (u >= 0.0f && u < m.w) {
        const int j = (int)u;
        const float t = u - (float)j;
        const float4* __restrict__ nd = a.atab + a.atab_off[i] + j;
        const float4 p0 = __ldg(nd), p1 = __ldg(nd + 1);
        const float t2 = t * t, t3 = t2 * t;
        const float h00 = 2.0f * t3 - 3.0f * t2 + 1.0f, h10 = (t3 - 2.0f * t2 + t) * m.z;
        const float h01 = 3.0f * t2 - 2.0f * t3, h11 = (t3 - t2) * m.z;
        sD = h00 * p0.x + h10 * p0.y + h01 * p1.x + h11 * p1.y;
        sK = h00 * p0.z + h10 * p0.w + h01 * p1.z + h11 * p1.w;
        done = true;
      }
    }
    if (!done) {
      const int2 rg = a.tab_range[i];
      const float4* __restrict__ tA = a.tabA + (size_t)i * a.limit;
      const float* __restrict__ tK = a.tabK + (size_t)i * a.limit;
      for (int n = rg.x; n < rg.y; ++n) {
        const float4 c = __ldg(tA + n);
        const float d1 = fmaf(k, c.x, c.y);
        const float d2 = d1 - c.z;
        sD = fmaf(c.w, ncdf(d1), sD);
        sK = fmaf(__ldg(tK + n), ncdf(d2), sK);
      }
    }
    A = Ge * sD - sK;
    dAb = (D == 1) ? sD : sD * Ge * (1.0f / D);
  }
  __device__ static __forceinline__ float dA_k(float dAb, float Xk) { return (D == 1) ? dAb : dAb / Xk; }
  // The same closed form when the G threads of a group hold the same state (jump-scheme kernels, one path per group): the
  // terms of the Poisson series are dealt out over the group and the two partial sums reduced by `allsum2` - the serial chain
  // of up to `limit` normal CDF pairs per step becomes one or two.
  template <class AllSum2>
  __device__ static __forceinline__ void eval_A_group(const PricingArgs& a, int i, const float (&X)[D], float& A, float& dAb, int G, int g,
                                                      AllSum2 allsum2) {
    const float Gm = basket(X);
    const float Ge = (D == 1) ? Gm : Gm * a.qdisc[i];
    const float k = logf(Ge / a.K);
    const int2 rg = a.tab_range[i];
    const float4* __restrict__ tA = a.tabA + (size_t)i * a.limit;
    const float* __restrict__ tK = a.tabK + (size_t)i * a.limit;
    float sD = 0.0f, sK = 0.0f;
    for (int n = rg.x + g; n < rg.y; n += G) {
      const float4 c = __ldg(tA + n);
      const float d1 = fmaf(k, c.x, c.y);
      const float d2 = d1 - c.z;
      sD = fmaf(c.w, ncdf(d1), sD);
      sK = fmaf(__ldg(tK + n), ncdf(d2), sK);
    }
    allsum2(sD, sK);
    A = Ge * sD - sK;
    dAb = (D == 1) ? sD : sD * Ge * (1.0f / D);
  }
  // Same closed form for the tcgen05 kernels, split in two so that the table loads are in flight while the caller does
  // other work: the d logarithms of the geometric mean collapse into one per half of the product (d = 10: two MUFU.LG2
  // instead of ten library logf + one expf).
  struct AEval { float k, Ge, t; float4 m, p0, p1; bool tab; };
  __device__ static __forceinline__ void eval_A_begin(const PricingArgs& a, int i, const float (&X)[D], AEval& e) {
    if (D == 1) {
      e.k = __logf(X[0] / a.K);
      e.Ge = X[0];
    } else {
      constexpr int D2 = D / 2;
      float p0 = 1.0f, p1 = 1.0f;
#pragma unroll
      for (int q = 0; q < D2; ++q) p0 *= X[q];
#pragma unroll
      for (int q = D2; q < D; ++q) p1 *= X[q];
      e.k = (__logf(p0) + __logf(p1)) * (1.0f / D) + __logf(a.qdisc[i] / a.K);
      e.Ge = a.K * __expf(e.k);
    }
    e.tab = false;
    if (a.use_atab) {
      e.m = __ldg(a.atab_meta + i);
      const float u = (e.k - e.m.x) * e.m.y;
      if (u >= 0.0f && u < e.m.w) {
        const int j = (int)u;
        e.t = u - (float)j;
        const float4* __restrict__ nd = a.atab + a.atab_off[i] + j;
        e.p0 = __ldg(nd); e.p1 = __ldg(nd + 1);
        e.tab = true;
      }
    }
  }
  __device__ static __forceinline__ void eval_A_finish(const PricingArgs& a, int i, const AEval& e, float& A, float& dAb) {
    float sD = 0.0f, sK = 0.0f;
    if (e.tab) {
      const float t = e.t, t2 = t * t, t3 = t2 * t;
      const float h00 = 2.0f * t3 - 3.0f * t2 + 1.0f, h10 = (t3 - 2.0f * t2 + t) * e.m.z;
      const float h01 = 3.0f * t2 - 2.0f * t3, h11 = (t3 - t2) * e.m.z;
      sD = h00 * e.p0.x + h10 * e.p0.y + h01 * e.p1.x + h11 * e.p1.y;
      sK = h00 * e.p0.z + h10 * e.p0.w + h01 * e.p1.z + h11 * e.p1.w;
    } else {
      const int2 rg = a.tab_range[i];
      const float4* __restrict__ tA = a.tabA + (size_t)i * a.limit;
      const float* __restrict__ tK = a.tabK + (size_t)i * a.limit;
      for (int n = rg.x; n < rg.y; ++n) {
        const float4 c = __ldg(tA + n);
        const float d1 = fmaf(e.k, c.x, c.y);
        const float d2 = d1 - c.z;
        sD = fmaf(c.w, ncdf(d1), sD);
        sK = fmaf(__ldg(tK + n), ncdf(d2), sK);
      }
    }
    A = e.Ge * sD - sK;
    dAb = (D == 1) ? sD : sD * e.Ge * (1.0f / D);
  }
  // jump-row inputs (SolversJumpDiff.py:37-39, 99-100, 173-175) incl. the constant-1 feature
  template <int HP>
  __device__ static __forceinline__ void jump_input(const PricingArgs& a, float t, const float (&X)[D],
                                                    const float (&Jv)[D], float (&in)[HP]) {
#pragma unroll
    for (int j = 0; j < HP; ++j) in[j] = 0.0f;
    in[0] = t;
    if (a.one_net) {
#pragma unroll
      for (int k = 0; k < D; ++k) in[1 + k] = X[k] * expf(Jv[k]);
      in[1 + D] = 1.0f;
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        in[1 + k] = X[k];
        in[1 + D + k] = a.feat_mode == 0 ? Jv[k] : expf(Jv[k]);
      }
      in[1 + 2 * D] = 1.0f;
    }
  }
  template <int HP>
  __device__ static __forceinline__ void jump_input_grad(const PricingArgs& a, const float (&Jv)[D], const float (&dx)[HP],
                                                         float (&dX)[D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) dX[k] += a.one_net ? dx[1 + k] * expf(Jv[k]) : dx[1 + k];
  }
  // Two-network schemes: the first layer of the jump network is SEPARABLE, W1 in = (time, state, 1 part) + scale * (sample part):
  // the D jump features sit in the input slots [kJumpSlot0, + D), feature k = jump_feature(J_k); scale = 1.
  static constexpr int kJumpSlots = D;
  __device__ static __forceinline__ int jump_slot0() { return 1 + D; }
  __device__ static __forceinline__ float jump_feature(const PricingArgs& a, float J) { return a.feat_mode == 0 ? J : expf(J); }
  __device__ static __forceinline__ float jump_scale(const float (&)[D]) { return 1.0f; }
  static constexpr bool kScaleIsState = false;
};

struct VGModel {
  static constexpr int D = 1;
  static constexpr bool kBrownian = false;

  __device__ static __forceinline__ float basket(const float (&X)[1]) { return X[0]; }
  // Lewis/FFT price through the per-step cubic table (pricingModels.py:156-179); the spline value is a
  // constant for the gradient (tf.numpy_function), only the explicit X - sqrt(X K) factors are differentiated.
  __device__ static __forceinline__ void eval_A(const PricingArgs& a, int i, const float (&X)[1], float& A, float& dAb) {
    const float x = X[0];
    const float k = logf(x / a.K);
    int idx = (int)floorf((k - a.vg_k0) * a.vg_inv_h);
    idx = idx < 0 ? 0 : (idx > a.vg_nint - 1 ? a.vg_nint - 1 : idx);
    const float t = k - (a.vg_k0 + (float)idx * a.vg_h);
    const float4 c = __ldg(a.vg_coef + (size_t)i * a.vg_nint + idx);
    const float S = fmaf(fmaf(fmaf(c.w, t, c.z), t, c.y), t, c.x);
    const float cs = a.vg_scale[i] * S;
    const float sq = sqrtf(x * a.K);
    A = x - sq * cs;
    dAb = 1.0f - 0.5f * sq / x * cs;
  }
  __device__ static __forceinline__ float dA_k(float dAb, float) { return dAb; }
  struct AEval { float A, dAb; };
  __device__ static __forceinline__ void eval_A_begin(const PricingArgs& a, int i, const float (&X)[1], AEval& e) {
    eval_A(a, i, X, e.A, e.dAb);
  }
  __device__ static __forceinline__ void eval_A_finish(const PricingArgs&, int, const AEval& e, float& A, float& dAb) {
    A = e.A; dAb = e.dAb;
  }
  // jump-row inputs (SolversPureJump.py:34-36, 95-96) incl. the constant-1 feature
  template <int HP>
  __device__ static __forceinline__ void jump_input(const PricingArgs& a, float t, const float (&X)[1],
                                                    const float (&Jv)[1], float (&in)[HP]) {
#pragma unroll
    for (int j = 0; j < HP; ++j) in[j] = 0.0f;
    in[0] = t;
    if (a.one_net) {
      in[1] = X[0] + X[0] * Jv[0];
      in[2] = 1.0f;
    } else {
      in[1] = X[0];
      in[2] = X[0] * Jv[0];
      in[3] = 1.0f;
    }
  }
  template <int HP>
  __device__ static __forceinline__ void jump_input_grad(const PricingArgs& a, const float (&Jv)[1], const float (&dx)[HP],
                                                         float (&dX)[1]) {
    if (a.one_net) dX[0] += dx[1] * (1.0f + Jv[0]);
    else dX[0] += dx[1] + dx[2] * Jv[0];
  }
  // separable first layer (two-network schemes): the jump feature X J = scale * J with scale = X (slot 2)
  static constexpr int kJumpSlots = 1;
  __device__ static __forceinline__ int jump_slot0() { return 2; }
  __device__ static __forceinline__ float jump_feature(const PricingArgs&, float J) { return J; }
  __device__ static __forceinline__ float jump_scale(const float (&X)[1]) { return X[0]; }
  static constexpr bool kScaleIsState = true;      // d(scale c)/dX = c: the adjoint adds sum_m d1_m . c_m to dL/dX
};

}  // namespace fbsdej

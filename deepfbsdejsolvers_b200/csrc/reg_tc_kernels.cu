// Adjoint sweep of the compensator-free (`Reg`) solvers with EVERY matrix product of the network on tcgen05:
//
//   SolverGlobalSumLocalReg  coupledPricing/SolversJumpDiff.py:391-415, SolversPureJump.py:361-384
//   SolverGlobalMultiStepReg coupledPricing/SolversJumpDiff.py:461-481, SolversPureJump.py:430-450
//   (their tf.GradientTape pass, SolversJumpDiff.py:421-427)
//
// One CTA = one tile of 128 paths, thread r = path r = TMEM lane r; the CTA walks the N time steps backwards.  Per step the
// network u(t, X) (nin = 1 + D -> H -> H -> 1) is re-evaluated and differentiated with six GEMMs, all bf16x3
// (x = hi + lo, D += A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM, ~1e-5 relative):
//
//   L1   acc = X  W1          K = 16   (inputs incl. the constant-1 feature that carries b1)
//   L2   acc = H1 W2          K = 32   (H1 incl. the constant-1 feature that carries b2)
//   WG2  [dW2 | dW3] += [H1 | H2]^T D2          rows are K (MN-major operands); column 23 of D2 holds dL/dy
//   BT   acc = D2 W2^T        K = 32
//   DX   acc = D1 W1^T        K = 32   (input gradient, feeds the adjoint of X)
//   WG1  dW1 += X^T D1
//
// Only the element-wise work stays on the CUDA cores (tanh, the deltas, the bf16 hi/lo split of the operand tiles and
// the adjoint of the coupled Euler step).  The weight-gradient accumulators live in TMEM for the whole kernel (all
// steps, all tiles of the CTA) and are read once at the end.  Each GEMM is issued by a different warp's lane 0, so
// the descriptor arithmetic is spread over the four warps.
//
// Shared memory (54.9 KB -> 4 CTAs per SM; 128 TMEM columns each):
//   operand tiles [feature / 8][128 rows][8 bf16] (16-byte chunks; one byte layout is K-major when features are K and
//   MN-major when rows are K, tc.cuh), hi and lo copies:  X (16 features), H1, H2, D2 (24 features); D1 reuses H2's
//   tile (its last reader, WG2, has completed when BT's commit is observed).
//   B operands of the K-major GEMMs ([k / 8][n][8 bf16], 24 n-rows per chunk): W1, W2, W2^T, W1^T, hi and lo.
//
// Per path-step inputs come from the tile-major record written by the forward sweep (pricing.cuh: RecLayout): one base
// pointer, immediate offsets, one bulk L2 prefetch per (tile, step).
#ifndef FBSDEJ_ABLATE
#define FBSDEJ_ABLATE 0   // timing experiments only (scripts/ablate_forward.sh): 1 no RNG math, 2 no coupling, 3 no tanh, 4 no MMA, 5 no X stores
#endif
#include "pricing.cuh"
#include "sim_device.cuh"
#include "tc.cuh"
#include "tc_net.cuh"

namespace fbsdej {
namespace rtc {

namespace bwd {
constexpr int CH = 128;                       // uint4 per chunk (128 rows x 16 bytes)
constexpr int XA_HI = 0, XA_LO = 2 * CH, H1_HI = 4 * CH, H2_HI = 7 * CH, H1_LO = 10 * CH, H2_LO = 13 * CH, D2_HI = 16 * CH,
              D2_LO = 19 * CH, D1_HI = H2_HI, D1_LO = H1_LO, W_BASE = 22 * CH;
// B operands of the layer GEMMs, hi and lo copies STACKED ALONG N inside every K chunk: [k / 8][n' ][8 bf16] with
// n' = n (hi) for n' < NH and n' - NH (lo) above; NH = 24 (W1, W2, W2^T) or 16 (W1^T)
constexpr int W1B = W_BASE, W2B = W1B + 2 * 2 * NB, WTB = W2B + 4 * 2 * NB, W1T = WTB + 4 * 2 * NB, U4_END = W1T + 4 * 2 * 16;
constexpr int OFF_W3 = U4_END * 4;            // float offsets after the uint4 region
constexpr int OFF_BAR = OFF_W3 + 24;          // two mbarriers (8-byte aligned) + the TMEM base slot
constexpr int SMEM_FLOATS = OFF_BAR + 8;
static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
constexpr uint32_t WIMG_BYTES = (uint32_t)(OFF_W3 + 24 - W_BASE * 4) * 4;   // weight operand image: the uint4 B operands + W3
static_assert(WIMG_BYTES % 16 == 0, "bulk copy size");
constexpr uint32_t C_ACC = 0, C_W1 = 48, C_W2 = 80, NCOLS = 128;   // ACC 48 | WG1 32 | WG2 48 columns
constexpr int COL_DOUT = 23;                  // spare column of D2 that carries dL/dy through WG2 (needs H <= 22)
// tanh layers in "r form" (as the forward sweep): the GEMM delivers x' = 2 log2(e) x (scale in the staged weights), the thread forms
// r = 1 / (2^x' + 1) and stores r - not h = 1 - 2 r - in the operand tiles; the next layer's weights carry the factor -2 and its
// bias row b + sum_k W[k][.]; 1 - h^2 = 4 (r - r^2) with the 4 in the staged W3 / W2^T.  Two instructions per hidden unit less
// than tanh_fast + dact.  The weight gradients come out against r: dW[k][j] = db[j] - 2 sum_rows r[k] d[j] (flush).
template <int ACT>
__device__ __forceinline__ float hid_r(float x) {
  if (ACT != ACT_TANH) return fmaxf(x, 0.0f);
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return r;
}
template <int ACT>
__device__ __forceinline__ float dact_r(float r) { return ACT == ACT_TANH ? fmaf(-r, r, r) : (r > 0.0f ? 1.0f : 0.0f); }
}  // namespace bwd

template <class Model, int ACT>
__global__ void __launch_bounds__(kThreads, 4) reg_backward_tc(const PricingArgs a) {
  constexpr int D = Model::D;
  using RL = RecLayout<D>;
  using namespace bwd;
  static_assert(D + 2 <= 16, "the X tile holds 16 features");
  extern __shared__ __align__(1024) float smem[];
  uint4* const u4 = reinterpret_cast<uint4*>(smem);
  float* const w3s = smem + OFF_W3;
  uint64_t* const bar_f = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* const bar_w = bar_f + 1;
  uint64_t* const bar_g = bar_f + 3;                    // (slot 2 holds the TMEM base)
  uint32_t* const tslot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 4);
  const int row = threadIdx.x, warp = row >> 5;
  const int H = a.netA.H, nin = a.netA.nin;

  // ---- one-time set-up: zero the tiles, TMA bulk copy of the weight operand image (bf16 hi / lo B operands + W3; built from
  // theta by reg_stage_operands_kernel), TMEM, barriers ------------------------------------------------------------------
  for (int i = row; i < SMEM_FLOATS; i += kThreads)
    if (i < W_BASE * 4 || i >= OFF_W3 + 24) smem[i] = 0.0f;
  __syncthreads();
  if (row == 0) { tc::mbar_init(bar_f, 1); tc::mbar_init(bar_w, 1); tc::mbar_init(bar_g, 1); tc::fence_mbar_init(); }
  __syncthreads();
  if (row == 0) {
    tc::mbar_arrive_expect_tx(bar_w, WIMG_BYTES);
    tc::bulk_g2s(u4 + W_BASE, a.wimg_bwd, WIMG_BYTES, bar_w);
  }
  tc::mbar_wait(bar_w, 0);                              // (phase 0 of bar_w; the WG1 commits use the following phases)
  if (warp == 0) tc::tmem_alloc(tslot, NCOLS);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16);
  // thread j <= H owns the effective layer-1 bias of hidden unit j: c_j(t) = t W1[0][j] + b1[j] (fp32, then split)
  float w0 = 0.0f, b1v = 0.0f;
  constexpr float CS = ACT == ACT_TANH ? 2.885390081777927f : 1.0f;        // pre-activation scale 2 log2(e) (r form)
  if (row < H) { w0 = CS * a.theta[a.netA.ext_off + row]; b1v = CS * a.theta[a.netA.ext_off + nin * H + row]; }
  const int bias_idx = ((nin >> 3) * 2 * NB + row) * 8 + (nin & 7);   // hi copy; the lo copy is NB n-rows further
  const uint32_t sbase = tc::smem_u32(u4);
  const uint32_t sbase16 = tc::addr16(sbase);                 // operand addresses in units of 16 bytes (tc::smem_desc16)
  auto sa = [&](int off_u4) { return sbase16 + (uint32_t)off_u4; };
  uint32_t phase_f = 0, phase_w = 1, phase_g = 0, pending_w = 0, started = 0;
  auto wait_f = [&]() { tc::mbar_wait(bar_f, phase_f); phase_f ^= 1; tc::tc_fence_after(); };

  const float invB = a.inv_B, invBN = a.inv_B / (float)a.N, rdt = a.r * a.dt;
  const int ntiles = a.tmap.ntiles;
  const uint32_t step_bytes = (uint32_t)(RL::NP * TR * sizeof(float));
  // rows beyond the tile's height (96-row tiles) were never written by the forward sweep: they run on benign constants
  const bool live = row < tile_height(a.tmap, blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float msk = (live && tile_base(a.tmap, tile) + row < a.B) ? 1.0f : 0.0f;
    const float* const rec0 = a.rec + (size_t)tile * a.N * RL::NP * TR + row;
    const float* const recN = a.recN + (size_t)tile * RL::NPT * TR + row;
    float Xbar[D];
    float Esum = 0.0f, rb_next = 0.0f;                  // MultiStep: running sum of e_k;  SumLocal: 2 rho_i / B of the step above
    {
      float X[D];
#pragma unroll
      for (int k = 0; k < D; ++k) X[k] = live ? recN[k * TR] : 1.0f;
      float gbar;
      if (a.scheme == SCH_MULTISTEP) {
        Esum = live ? recN[D * TR] : 0.0f;              // sum_k e_k ; d loss / d g = -2/(NB) sum_k e_k
        gbar = -2.0f * Esum * invBN;
      } else {
        rb_next = live ? 2.0f * rec0[((size_t)(a.N - 1) * RL::NP + RL::P_SCH) * TR] * invB : 0.0f;
        gbar = rb_next;
      }
      const float Gb = Model::basket(X);
      const float ind = (Gb - a.K >= 0.0f) ? 1.0f : 0.0f;   // tf.maximum: gradient to the first argument on ties
#pragma unroll
      for (int k = 0; k < D; ++k) Xbar[k] = gbar * ind * ((D == 1) ? 1.0f : Gb / ((float)D * X[k]));
    }
    if (row == 0 && a.N >= 2) prefetch_l2_bulk(rec0 + (size_t)(a.N - 2) * RL::NP * TR, step_bytes);
    // the record of step i is loaded one step ahead (during the last MMA wait of step i + 1): one base pointer,
    // immediate offsets, DRAM / L2 latency off the critical chain
    float Xn[D], En[D], s_n = 0.0f, dA_n = 0.0f, sch_n = 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) { Xn[k] = 1.0f; En[k] = 1.0f; }
    auto load_step = [&](int i) {
      const float* const rs = rec0 + (size_t)i * RL::NP * TR;
      if (live) {
#pragma unroll
        for (int k = 0; k < D; ++k) { Xn[k] = rs[(RL::P_X + k) * TR]; En[k] = rs[(RL::P_E + k) * TR]; }
        s_n = rs[RL::P_S * TR]; dA_n = rs[RL::P_DA * TR];
        // MultiStep: e_i ; SumLocal: rho_{i-1} (the record of the step below; for i = 0 the value is unused)
        sch_n = rs[(a.scheme == SCH_MULTISTEP || i == 0) ? RL::P_SCH * TR : (RL::P_SCH - RL::NP) * TR];
      }
      if (row == 0 && i >= 3) prefetch_l2_bulk(rs - 3 * RL::NP * TR, step_bytes);
    };
    load_step(a.N - 1);
    for (int i = a.N - 1; i >= 0; --i) {
      float X[D], E[D];
#pragma unroll
      for (int k = 0; k < D; ++k) { X[k] = Xn[k]; E[k] = En[k]; }
      const float s_i = s_n, dAb = dA_n, sch = sch_n;
      // ---- adjoint of the coupled Euler step X' = X E + aLin |y - A(i, X)| dt and of the loss graph -----------
      float sumXbar = 0.0f;
#pragma unroll
      for (int k = 0; k < D; ++k) sumXbar += Xbar[k];
      const float cY = sumXbar * s_i;
      const float cA = cY * dAb;
#pragma unroll
      for (int k = 0; k < D; ++k) Xbar[k] = fmaf(Xbar[k], E[k], -((D == 1) ? cA : cA * rcp_fast(X[k])));
      float ybar;
      if (a.scheme == SCH_MULTISTEP) {
        const float abar = 2.0f * Esum * invBN;           // sum_{k<=i} Fbar_k
        ybar = 2.0f * sch * invBN + rdt * abar + cY;
        Esum -= sch;
      } else {
        const float rb = rb_next;
        const float rbm = (i > 0) ? 2.0f * sch * invB : 0.0f;
        ybar = rbm - rb - rdt * rb + cY;
        rb_next = rbm;
      }
      const float dout = ybar * msk;
      const float tf = (a.scheme == SCH_SUMLOCAL && a.stale_time) ? (float)(i == 0 ? 0 : i - 1) : (float)i;
      // ---- X tile (inputs incl. the constant 1) -> L1 -------------------------------------------------------------
      if (pending_w) { tc::mbar_wait(bar_w, phase_w); phase_w ^= 1; pending_w = 0; }   // WG1 of the step above read X
      {
        float xin[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) xin[k] = 0.0f;
        xin[0] = tf;
#pragma unroll
        for (int k = 0; k < D; ++k) xin[1 + k] = X[k];
        xin[1 + D] = 1.0f;
        tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, 0, row, xin);
        tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, 1, row, xin + 8);
        if (row <= H) {
          uint32_t hi, lo;
          tc::split_bf16(row < H ? fmaf(tf, w0, b1v) : (ACT == ACT_TANH ? -200.0f : 1.0f), hi, lo);   // constant unit: r(-200) = 1
          reinterpret_cast<unsigned short*>(u4 + W1B)[bias_idx] = (unsigned short)hi;
          reinterpret_cast<unsigned short*>(u4 + W1B)[bias_idx + NB * 8] = (unsigned short)lo;
        }
      }
      publish();
      if (warp == 0 && tc::elect_one()) {
        tc::tc_fence_after();
#if FBSDEJ_ABLATE != 12
        gemm_k16<1, NB>(tmem + C_ACC, sa(XA_HI), sa(XA_LO), sa(W1B));
#endif
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- h1 -> L2 ---------------------------------------------------------------------------------------------
      float h1[24];
      load_acc<NB, 24>(lane_base + C_ACC, h1);
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) h1[8 * c8 + q] = hid_r<ACT>(h1[8 * c8 + q]);
        tc::store_bf16x8(u4 + H1_HI, u4 + H1_LO, c8, row, h1 + 8 * c8);
      }
      publish();
      if (warp == 1 && tc::elect_one()) {
        tc::tc_fence_after();
#if FBSDEJ_ABLATE != 12
        gemm_k16<2, NB>(tmem + C_ACC, sa(H1_HI), sa(H1_LO), sa(W2B));
#endif
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- h2, delta 2 -> WG2, BT -------------------------------------------------------------------------------
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float t8[8], d2[8];
        {
          float q8[8];
          tc::tmem_ld8(lane_base + C_ACC + 8 * c8, t8);
          tc::tmem_ld8(lane_base + C_ACC + NB + 8 * c8, q8);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) t8[q] += q8[q];
        }
        const float4 wa = ld4(w3s + 8 * c8), wb = ld4(w3s + 8 * c8 + 4);
        const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float h = hid_r<ACT>(t8[q]);
          t8[q] = h;
          d2[q] = dout * w8[q] * dact_r<ACT>(h);
        }
        if (c8 == 2) d2[COL_DOUT - 16] = dout;
        tc::store_bf16x8(u4 + H2_HI, u4 + H2_LO, c8, row, t8);
        tc::store_bf16x8(u4 + D2_HI, u4 + D2_LO, c8, row, d2);
      }
      publish();
      if (warp == 2 && tc::elect_one()) {
        tc::tc_fence_after();
        // the input-gradient GEMM first: its result is on the step's critical chain, the weight-gradient GEMM is not (the
        // tensor pipe is in order); the weight-gradient GEMM gets its own barrier because D1 overwrites tiles it reads
#if FBSDEJ_ABLATE != 12
        gemm_k16<2, NB>(tmem + C_ACC, sa(D2_HI), sa(D2_LO), sa(WTB));
#endif
        tc::mma_commit(bar_f);
#if FBSDEJ_ABLATE != 11
        gemm_rows_stacked16<48>(tmem + C_W2, sa(H1_HI), sa(D2_HI), started ? 1u : 0u);   // [H1 | H2 (hi) | H1 | H2 (lo)]^T [D2 hi | lo]
#endif
        tc::mma_commit(bar_g);
      }
      wait_f();
      // ---- delta 1 -> DX, WG1 ---------------------------------------------------------------------------------
      {
        float d1[24];
#pragma unroll
        for (int c8 = 0; c8 < 3; ++c8) {
          float q8[8];
          tc::tmem_ld8(lane_base + C_ACC + 8 * c8, reinterpret_cast<float (&)[8]>(d1[8 * c8]));
          tc::tmem_ld8(lane_base + C_ACC + NB + 8 * c8, q8);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) d1[8 * c8 + q] = (d1[8 * c8 + q] + q8[q]) * dact_r<ACT>(h1[8 * c8 + q]);
        }
        tc::mbar_wait(bar_g, phase_g); phase_g ^= 1;     // WG2 has read H2 / H1_lo: their tiles may become D1
#pragma unroll
        for (int c8 = 0; c8 < 3; ++c8) tc::store_bf16x8(u4 + D1_HI, u4 + D1_LO, c8, row, d1 + 8 * c8);
      }
      publish();
      if (warp == 3 && tc::elect_one()) {
        tc::tc_fence_after();
#if FBSDEJ_ABLATE != 12
        gemm_k16<2, 16>(tmem + C_ACC, sa(D1_HI), sa(D1_LO), sa(W1T));
#endif
        tc::mma_commit(bar_f);
#if FBSDEJ_ABLATE != 11
        gemm_rows_stacked16<32, 64>(tmem + C_W1, sa(D1_HI), sa(XA_HI), started ? 1u : 0u);   // [D1 hi | D1 lo]^T [X hi | X lo] = dW1^T (M = 64)
#endif
        tc::mma_commit(bar_w);
      }
      started = 1;
      pending_w = 1;
      if (i > 0) load_step(i - 1);
      wait_f();
      {
        float dx[16];
        load_acc<16, (D > 7 ? 16 : 8)>(lane_base + C_ACC, reinterpret_cast<float (&)[D > 7 ? 16 : 8]>(dx));
#pragma unroll
        for (int k = 0; k < D; ++k) Xbar[k] += dx[1 + k];
      }
      tc::tc_fence_before();     // orders these TMEM reads before the next step's first MMA (via its publish barrier)
    }
  }
  // ---- flush: TMEM weight gradients -> smem vector (external flat layout) -> this CTA's row of gpart ------------
  if (pending_w) { tc::mbar_wait(bar_w, phase_w); phase_w ^= 1; pending_w = 0; }
  tc::tc_fence_after();
  __syncthreads();
  float* const S = smem;                               // the operand tiles are dead: [128 lanes][49] scratch
  float* const sg = smem + 8192;                       // gradient vector, external flat layout
  constexpr int SW = 49;
  for (int e = row; e < a.P; e += kThreads) sg[e] = 0.0f;
  float* const g = sg + a.netA.ext_off;
  const int o2 = nin * H + H, o3 = o2 + H * H + H;
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
    if (started) {
#pragma unroll
      for (int c8 = 0; c8 < (pass == 0 ? 4 : 6); ++c8) {
        float v[8];
        tc::tmem_ld8(lane_base + (pass == 0 ? C_W1 : C_W2) + 8 * c8, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) S[row * SW + 8 * c8 + q] = v[q];
      }
    }
    __syncthreads();
    if (started && pass == 0) {                        // rows (M = 64 lane map): D1 hi j, D1 lo 24 + j; columns: X hi i, X lo 16 + i
      for (int e = row; e < (nin + 1) * H; e += kThreads) {
        const int i = e / H, j = e % H;                // i = nin: b1
        const int lh = lane_of_row_m64(j), ll = lane_of_row_m64(24 + j);
        g[e] = (S[lh * SW + i] + S[lh * SW + 16 + i]) + (S[ll * SW + i] + S[ll * SW + 16 + i]);   // hi.hi + hi.lo + lo.hi + lo.lo
      }
    } else if (started) {                              // lanes: H1 hi 0..23, H2 hi 24..47, H1 lo 48..71, H2 lo 72..95
      auto w2sum = [&](int k, int j) { return (S[k * SW + j] + S[k * SW + 24 + j]) + (S[(48 + k) * SW + j] + S[(48 + k) * SW + 24 + j]); };
      auto w3sum = [&](int k) {
        return (S[(24 + k) * SW + COL_DOUT] + S[(24 + k) * SW + 24 + COL_DOUT]) + (S[(72 + k) * SW + COL_DOUT] + S[(72 + k) * SW + 24 + COL_DOUT]);
      };
      // tanh: the tiles hold r = (1 - h) / 2, so sum_rows h[k] d[j] = db[j] - 2 sum_rows r[k] d[j] (the constant unit, k = H, is r = 1)
      for (int e = row; e < (H + 1) * H; e += kThreads) {
        const int k = e / H, j = e % H;                // k = H: b2
        g[o2 + e] = (ACT == ACT_TANH && k < H) ? fmaf(-2.0f, w2sum(k, j), w2sum(H, j)) : w2sum(k, j);
      }
      if (row <= H)                                    // dW3[k] (k = H: b3) rides in column COL_DOUT of D2
        g[o3 + row] = (ACT == ACT_TANH && row < H) ? fmaf(-2.0f, w3sum(row), w3sum(H)) : w3sum(row);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  float* const grow = a.gpart + (size_t)blockIdx.x * a.P;
  for (int e = row; e < a.P; e += kThreads) grow[e] = sg[e];
  if (warp == 0) tc::tmem_dealloc(tmem, NCOLS);
}

// ---- forward sweep ---------------------------------------------------------------------------------------------
// Loss graphs: SolverGlobalSumLocalReg.regressOptim (SolversJumpDiff.py:391-415), SolverGlobalMultiStepReg.regressOptim
// (:461-481) and their VG twins; model step pricingModels.py:53-54 / :184-185.
// Both layers run on tcgen05 with the 3xTF32 split (fp32-grade: the loss and the stored trajectories keep 1e-5 parity);
// the time feature is folded into a per-step effective bias c_j = t W1[0][j] + b1[j] formed in fp32 (it is uniform over
// the tile), so no operand of the split GEMMs is larger than O(1).
//
// One CTA = one tile of 128 paths = TWO warpgroups with different jobs (an SM holds 4 such CTAs = 32 warps; with one thread per
// path and 443 paths per SM at 2^16 paths there were 14, and the sweep was latency-bound):
//   warpgroup 0 ("network"): thread r = path r = TMEM lane r.  Operand writes to tensor memory, the two MMA round trips, tanh,
//                the closed-form coupling A(i, X), the loss-graph bookkeeping, the coupled Euler step and the record stores.
//   warpgroup 1 ("increments"): thread r produces path r's exponentials E = e^{drift dt + sig dW + J} one to two steps ahead -
//                drawn from Philox counters (RNG; sim_device.cuh) or read from the materialised dW / J planes - into a
//                two-stage shared-memory ring (mbarrier full / empty pairs, one arrival per warp) and into the record.
// The increments do not depend on the state, so the two chains only meet at the ring; setmaxnreg moves registers from the
// producer to the network warpgroup.
namespace fwd {
constexpr int NST = 2;                        // ring stages
// shared memory: the B operands (weights) - the activations go to tensor memory - and the increment ring
constexpr int W1B_HI = 0, W1B_LO = W1B_HI + 4 * NB * 4, W2B_HI = W1B_LO + 4 * NB * 4, W2B_LO = W2B_HI + 6 * NB * 4,
              OFF_W3 = W2B_LO + 6 * NB * 4 + 32 /* N = 32 reads 8 n-rows past the last chunk */, OFF_RED = OFF_W3 + 32,
              OFF_BAR = OFF_RED + 8 /* mbarriers: mma, full[NST], empty[NST] */, OFF_TS = OFF_BAR + 2 * (1 + 2 * NST),
              OFF_THR = OFF_TS + 2 /* 64 Poisson thresholds (fused RNG) */, OFF_RING = OFF_THR + 64;
static_assert((OFF_BAR % 2) == 0 && (OFF_RING % 4) == 0 && (OFF_RED % 4) == 0, "mbarrier / ring / image alignment");
template <int D> constexpr int smem_floats() { return OFF_RING + NST * D * TR; }
// tensor memory: two allocations (32 + 64 = 96 columns): the accumulator, and the A operand hi (X: 16, H1: 24 columns) | lo
constexpr uint32_t NCOLS_ACC = 32, NCOLS_A = 64;
constexpr int REGS_NET = 80, REGS_INC = 48;   // (80 + 48) * 128 threads = 64 * 256: four CTAs per SM

// hidden unit in r form (tanh: r = 1 / (2^x' + 1), h = 1 - 2 r is never formed) or ReLU
template <int ACT>
__device__ __forceinline__ float hidf(float x) {
  if (ACT != ACT_TANH) return fmaxf(x, 0.0f);
#if FBSDEJ_ABLATE == 3
  return 0.5f - 0.1f * x;
#endif
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return r;
}
__device__ __forceinline__ void net_barrier(int nthr) { asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); }   // warpgroup 0 only
// TMEM operand writes (+ the bias row in shared memory) -> visible to the MMA issued after the warpgroup barrier
__device__ __forceinline__ void publish_net(int nthr) {
  tc::tmem_st_wait();
  tc::fence_async_smem();
  tc::tc_fence_before();
  net_barrier(nthr);
}
}  // namespace fwd

template <class Model, int ACT, bool RNG>
__global__ void __launch_bounds__(2 * kThreads, 4) reg_forward_tc(const PricingArgs a) {
  constexpr int D = Model::D;
  using RL = RecLayout<D>;
  using namespace fwd;
  static_assert(D + 2 <= 16, "the X tile holds 16 features");
  extern __shared__ __align__(1024) float smem[];
  float* const w3s = smem + OFF_W3;
  float* const red = smem + OFF_RED;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [0] MMA, [1 + s] full, [1 + NST + s] empty
  uint32_t* const tslot = reinterpret_cast<uint32_t*>(smem + OFF_TS);
  float* const ring = smem + OFF_RING;                                 // [NST][D][128]
  const int tid = threadIdx.x, row = tid & (kThreads - 1), wg = tid >> 7, warp = (tid >> 5) & 3, lane = tid & 31;
  const int H = a.netA.H, nin = a.netA.nin;
  constexpr float CS = ACT == ACT_TANH ? 2.885390081777927f : 1.0f;        // pre-activation scale 2 log2(e)
  constexpr float one_in = ACT == ACT_TANH ? -200.0f : 1.0f;               // constant-1 unit: r(-200) = 1, relu(1) = 1

  // weight operand image (TF32 hi / lo B operands in r form + W3; reg_stage_operands_kernel) by one TMA bulk copy
  for (int i = tid + OFF_RED; i < OFF_RING; i += 2 * kThreads) smem[i] = 0.0f;
  __syncthreads();
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(bar, OFF_RED * 4);
    tc::bulk_g2s(smem, a.wimg_fwd, OFF_RED * 4, bar);
  }
  tc::mbar_wait(bar, 0);                                 // (phase 0 of the MMA barrier; the commits use the following phases)
  uint32_t* const sthr = reinterpret_cast<uint32_t*>(smem + OFF_THR);
  if (RNG && tid < 64) sthr[tid] = tid < a.npois ? a.pois_thr[tid] : 0xffffffffu;
  if (tid < 32) { tc::tmem_alloc(tslot, NCOLS_ACC, false); tc::tmem_alloc(tslot + 1, NCOLS_A); }
  const int hrows = tile_height(a.tmap, blockIdx.x);    // rows (= threads per warpgroup) of this CTA's tiles: 128 or 96
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { tc::mbar_init(bar + 1 + s, hrows / 32); tc::mbar_init(bar + 1 + NST + s, hrows / 32); }
    tc::fence_mbar_init();
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  const size_t sB = (size_t)a.B;
  const int ntiles = a.tmap.ntiles;
  uint32_t it = 0;                                      // ring position: steps of all tiles of this CTA, in order
  if (row >= hrows) return;                             // 96-row tiles: the fourth warp of either warpgroup has no paths

  if (wg == 1) {
    // ================================ increments: producer of the E ring ================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_INC));
    const uint32_t rng_iter = RNG ? (a.iter_ptr ? *a.iter_ptr : a.iteration) : 0u;
    const uint32_t t0 = RNG ? sthr[0] : 0u, t1 = RNG ? sthr[1] : 0u;
    const float inv_w1 = t1 > t0 ? 1.0f / (float)(t1 - t0) : 0.0f;
    // E = 2^(log2(e) (drift dt + sig dW + J)): log2(e) rides in the constants (jump sizes are linear in muJ, sigJ)
    constexpr float L2E = 1.4426950408889634f;
    const float driftL = L2E * a.drift_dt, sigL = L2E * a.sig, muJL = L2E * a.muJ, sigJL = L2E * a.sigJ;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int p0 = tile_base(a.tmap, tile) + row;
      const int p = p0 < a.B ? p0 : a.B - 1;
      const uint32_t gid = a.path_offset + (uint32_t)p;
      float* const rec0 = a.rec + (size_t)tile * a.N * RL::NP * TR + row;
      for (int i = 0; i < a.N; ++i, ++it) {
        const uint32_t s = it % NST;
        float E[D];
#if FBSDEJ_ABLATE == 1
        if constexpr (false) {
#else
        if constexpr (RNG) {
#endif
          // one Philox block per asset pair.  Jump sizes: a draw with exactly one jump takes its size from the residual of the
          // Poisson inversion through the inverse-normal polynomial (sim_device.cuh: jump_sizes_fast - same classification, same
          // arithmetic, hence the same bits as sim_merton_kernel).  At lam dt = 0.03 a path-step has such a draw with probability
          // 0.26 and two of them with 0.03, so the polynomial is evaluated ONCE per step, on the pending candidate, instead of
          // once per pair; a second candidate flushes the pending one first (divergent, rare).  Rare cells (a count >= 2, both
          // draws of a pair jumping, a far-tail size: ~0.4 %) are redone on the exact path below.
          constexpr int KP = (D + 1) / 2;
          float jj[2 * KP];
#pragma unroll
          for (int k = 0; k < 2 * KP; ++k) jj[k] = 0.0f;
          uint32_t rare_any = 0, cand_u = t0, cand_k = 255u;
          auto flush_candidate = [&]() {
            const float v = ((float)(cand_u - t0) + 0.5f) * inv_w1;
            const float x = fmaf(2.0f, v, -1.0f);
            float w = -0.6931471805599453f * lg2_raw(fmaf(-x, x, 1.0f));
            const bool tail = !(w < 5.0f);
            w -= 2.5f;
            float q = 2.81022636e-08f;
            q = fmaf(q, w, 3.43273939e-07f);
            q = fmaf(q, w, -3.5233877e-06f);
            q = fmaf(q, w, -4.39150654e-06f);
            q = fmaf(q, w, 0.00021858087f);
            q = fmaf(q, w, -0.00125372503f);
            q = fmaf(q, w, -0.00417768164f);
            q = fmaf(q, w, 0.246640727f);
            q = fmaf(q, w, 1.50140941f);
            const float jump = fmaf(sigJL, 1.4142135623730951f * q * x, muJL);
            if (tail && cand_k != 255u) rare_any |= 1u << cand_k;       // far tail: exact path
#pragma unroll
            for (int k = 0; k < 2 * KP; ++k) jj[k] = (cand_k == (uint32_t)k) ? jump : jj[k];
          };
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const uint4 r = Philox::rand4(gid, ((uint32_t)i << 8) | (uint32_t)kp, rng_iter, STREAM_PATH, a.seed_lo, a.seed_hi);
            float w0, w1;
            box_muller_fast(r.x, r.y, a.sqdt, w0, w1);
            const bool one0 = (r.z >= t0) && (r.z < t1), one1 = (r.w >= t0) && (r.w < t1);
            const bool both = one0 && one1;
            rare_any |= (((r.z >= t1) || both) ? 1u : 0u) << (2 * kp) | (((r.w >= t1) || both) ? 2u : 0u) << (2 * kp);
            if (one0 != one1) {                            // exactly one single jump in this pair
              if (cand_k != 255u) flush_candidate();
              cand_u = one1 ? r.w : r.z;
              cand_k = (uint32_t)(2 * kp) + (one1 ? 1u : 0u);
            }
            E[2 * kp] = fmaf(sigL, w0, driftL);
            if (2 * kp + 1 < D) E[2 * kp + 1] = fmaf(sigL, w1, driftL);
          }
          flush_candidate();
          if (rare_any) {
#pragma unroll
            for (int kp = 0; kp < KP; ++kp) {
              if ((rare_any >> (2 * kp)) & 3u) {
                const uint32_t c1 = ((uint32_t)i << 8) | (uint32_t)kp;
                const uint4 r = Philox::rand4(gid, c1, rng_iter, STREAM_PATH, a.seed_lo, a.seed_hi);
                if ((rare_any >> (2 * kp)) & 1u)
                  jj[2 * kp] = jump_size_rare(r.z, t0, t1, inv_w1, sthr, a.npois, muJL, sigJL, gid, c1, rng_iter, STREAM_PATH, a.seed_lo, a.seed_hi, 0);
                if ((rare_any >> (2 * kp)) & 2u)
                  jj[2 * kp + 1] = jump_size_rare(r.w, t0, t1, inv_w1, sthr, a.npois, muJL, sigJL, gid, c1, rng_iter, STREAM_PATH, a.seed_lo, a.seed_hi, 1);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < D; ++k) E[k] = ex2_raw(E[k] + jj[k]);
#if FBSDEJ_ABLATE == 1
        } else if (true) {
#pragma unroll
          for (int k = 0; k < D; ++k) E[k] = 1.0f + 1e-3f * (float)k;
#endif
        } else {
          const float* __restrict__ pw = a.dW + (size_t)i * D * sB + p;
          const float* __restrict__ pj = a.J + (size_t)i * D * sB + p;
#pragma unroll
          for (int k = 0; k < D; ++k)
            E[k] = __expf(a.drift_dt + (Model::kBrownian ? a.sig * pw[(size_t)k * sB] : 0.0f) + pj[(size_t)k * sB]);
        }
        if (it >= NST) tc::mbar_wait(bar + 1 + NST + s, ((it / NST) - 1) & 1);   // the network has read this stage
        float* const rg = ring + (size_t)s * D * TR + row;
        float* const rs = rec0 + (size_t)i * RL::NP * TR;
#pragma unroll
        for (int k = 0; k < D; ++k) { rg[k * TR] = E[k]; rs[(RL::P_E + k) * TR] = E[k]; }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar + 1 + s);
      }
    }
    return;
  }

  // ==================================== network: consumer of the ring ==================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_NET));
  const bool issuer = lane == 0;
  float w0 = 0.0f, b1v = 0.0f;                          // thread j <= H owns the effective bias of hidden unit j
  if (row < H) { w0 = CS * a.theta[a.netA.ext_off + row]; b1v = CS * a.theta[a.netA.ext_off + nin * H + row]; }
  const uint32_t tmem = tslot[0], tmem_a = tslot[1];
  const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16), lane_a = tmem_a + ((uint32_t)(row & ~31) << 16);
  const uint32_t sbase = tc::smem_u32(smem);
  const uint32_t sbase16 = tc::addr16(sbase);                 // operand addresses in units of 16 bytes (tc::smem_desc16)
  auto sa = [&](int off_f) { return sbase16 + (uint32_t)(off_f >> 2); };
  uint32_t phase = 1;
  auto wait_mma = [&]() { tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after(); };
  const int bias_idx = ((nin >> 2) * NB + row) * 4 + (nin & 3);   // W1B[n = row][k = nin]
  const float rdt = a.r * a.dt;
  float lsum = 0.0f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile_base(a.tmap, tile) + row;
    const bool valid = p0 < a.B;
    const int p = valid ? p0 : a.B - 1;
    float* const rec0 = a.rec + (size_t)tile * a.N * RL::NP * TR + row;
    float X[D];
#pragma unroll
    for (int k = 0; k < D; ++k) X[k] = a.x0;
    float Cpre = 0.0f;                               // MultiStep: sum_{j<i} toAdd_j
    float yprev = 0.0f, aprev = 0.0f, lloc = 0.0f;   // SumLocal
    for (int i = 0; i < a.N; ++i, ++it) {
      const float tf = (a.scheme == SCH_SUMLOCAL && a.stale_time) ? (float)(i == 0 ? 0 : i - 1) : (float)i;
      float* const rs = rec0 + (size_t)i * RL::NP * TR;
      {
        float xin[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) xin[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) xin[1 + k] = X[k];
        xin[1 + D] = 1.0f;
        if (row <= H) {
          float hi, lo;
          tc::split_tf32(row < H ? fmaf(tf, w0, b1v) : one_in, hi, lo);
          smem[W1B_HI + bias_idx] = hi;
          smem[W1B_LO + bias_idx] = lo;
        }
        // the TMEM stores go last, right before the wait::st of publish_net(): tcgen05.st reads its source registers
        // asynchronously, so any instruction that reuses one of them would stall until the store has drained
        store_tf32x8(lane_a, 0, xin);
        store_tf32x8(lane_a, 1, xin + 8);
      }
      publish_net(hrows);
      if (warp == 0 && tc::elect_one()) {
        tc::tc_fence_after();
#if FBSDEJ_ABLATE != 4
        gemm_k_tf32_16<2>(tmem, tmem_a, sa(W1B_HI), sa(W1B_LO));
#endif
        tc::mma_commit(bar);
      }
      // ---- independent of the network: closed-form coupling, record stores (in the shadow of the first MMA) ------------
#if FBSDEJ_ABLATE != 2
      typename Model::AEval ae;
      Model::eval_A_begin(a, i, X, ae);                  // table loads in flight until the shadow of the SECOND MMA (an L2 round
#endif                                                   // trip is longer than the first MMA: consuming them here stalled the chain)
#if FBSDEJ_ABLATE != 5
#pragma unroll
      for (int k = 0; k < D; ++k) rs[(RL::P_X + k) * TR] = X[k];
#endif
      wait_mma();
      {
        float t24[24];
        tc::tmem_ld8(lane_base, reinterpret_cast<float (&)[8]>(t24[0]));
        tc::tmem_ld8(lane_base + 8, reinterpret_cast<float (&)[8]>(t24[8]));
        tc::tmem_ld8(lane_base + 16, reinterpret_cast<float (&)[8]>(t24[16]));
        tc::tmem_ld_wait();                              // one drain for the three loads
#pragma unroll
        for (int q = 0; q < 24; ++q) t24[q] = hidf<ACT>(t24[q]);
        store_tf32x8(lane_a, 0, t24);                    // (L1 has completed: the X columns are free)
        store_tf32x8(lane_a, 1, t24 + 8);
        store_tf32x8(lane_a, 2, t24 + 16);
      }
      publish_net(hrows);
      if (warp == 1 && tc::elect_one()) {
        tc::tc_fence_after();
#if FBSDEJ_ABLATE != 4
        gemm_k_tf32_16<3>(tmem, tmem_a, sa(W2B_HI), sa(W2B_LO));
#endif
        tc::mma_commit(bar);
      }
      // ---- closed-form coupling, this step's exponentials from the ring (in the shadow of the second MMA) -----------------
#if FBSDEJ_ABLATE == 2
      float Ai = 0.2f, dAb = 0.5f;
#else
      float Ai, dAb;
      Model::eval_A_finish(a, i, ae, Ai, dAb);
#endif
      rs[RL::P_DA * TR] = dAb;
      float E[D];
      {
        const uint32_t s = it % NST;
        tc::mbar_wait(bar + 1 + s, (it / NST) & 1);
        const float* const rg = ring + (size_t)s * D * TR + row;
#pragma unroll
        for (int k = 0; k < D; ++k) E[k] = rg[k * TR];
        __syncwarp();
        if (issuer) tc::mbar_arrive(bar + 1 + NST + s);
      }
      wait_mma();
      float y_net = w3s[24];
      {
        float t24[24];
        tc::tmem_ld8(lane_base, reinterpret_cast<float (&)[8]>(t24[0]));
        tc::tmem_ld8(lane_base + 8, reinterpret_cast<float (&)[8]>(t24[8]));
        tc::tmem_ld8(lane_base + 16, reinterpret_cast<float (&)[8]>(t24[16]));
        tc::tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 6; ++c4) {
          const float4 w = ld4(w3s + 4 * c4);
          y_net = fmaf(hidf<ACT>(t24[4 * c4]), w.x, y_net);
          y_net = fmaf(hidf<ACT>(t24[4 * c4 + 1]), w.y, y_net);
          y_net = fmaf(hidf<ACT>(t24[4 * c4 + 2]), w.z, y_net);
          y_net = fmaf(hidf<ACT>(t24[4 * c4 + 3]), w.w, y_net);
        }
      }
      tc::tc_fence_before();
      // ---- loss-graph bookkeeping ---------------------------------------------------------------------------------
      if (a.trajY && valid) a.trajY[(size_t)i * sB + p] = y_net;
      const float ai = rdt * y_net;                        // "toAdd" = -dt f(Y)   (f = -r Y)
      if (a.scheme == SCH_MULTISTEP) {
        rs[RL::P_SCH * TR] = y_net - Cpre;                 // u_i ; F_i - g = u_i + (sum_all toAdd - g)
        Cpre += ai;
      } else {
        if (i > 0) {
          const float rho = y_net - yprev - aprev;
          lloc = fmaf(rho, rho, lloc);
          (rs - RL::NP * TR)[RL::P_SCH * TR] = rho;
        }
        yprev = y_net; aprev = ai;
      }
      // ---- coupled Euler step (pricingModels.py:53-54 / :184-185): the network's Y_i feeds the coupling ------------
      const float diff = y_net - Ai;
      const float coup = a.aLin * fabsf(diff) * a.dt;
      rs[RL::P_S * TR] = a.aLin * a.dt * (diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f));
#pragma unroll
      for (int k = 0; k < D; ++k) X[k] = fmaf(X[k], E[k], coup);
    }
    // ---- terminal condition -----------------------------------------------------------------------------------------
    const float gN = fmaxf(Model::basket(X) - a.K, 0.0f);
    float lpath;
    if (a.scheme == SCH_MULTISTEP) {
      // second sweep over the stored u_k: e_k = F_k - g(X_N), loss = mean_k mean_b e_k^2 (SolversJumpDiff.py:115)
      const float Dv = Cpre - gN;
      float se = 0.0f, s2 = 0.0f;
      for (int k = 0; k < a.N; ++k) {
        float* const q = rec0 + ((size_t)k * RL::NP + RL::P_SCH) * TR;
        const float e = *q + Dv;
        *q = e;
        se += e;
        s2 = fmaf(e, e, s2);
      }
      lpath = s2 * (a.inv_B / (float)a.N);
      a.recN[((size_t)tile * RL::NPT + D) * TR + row] = se;
    } else {
      const float rho = gN - yprev - aprev;
      lloc = fmaf(rho, rho, lloc);
      lpath = lloc * a.inv_B;
      rec0[((size_t)(a.N - 1) * RL::NP + RL::P_SCH) * TR] = rho;
    }
    if (a.trajY && valid) a.trajY[(size_t)a.N * sB + p] = gN;
#pragma unroll
    for (int k = 0; k < D; ++k) a.recN[((size_t)tile * RL::NPT + k) * TR + row] = X[k];
    if (valid) lsum += lpath;
  }
  tc::tc_fence_before();
  lsum = warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  net_barrier(hrows);
  if (row == 0) {
    a.lpart[blockIdx.x * 4] = red[0] + red[1] + red[2] + red[3];   // (red[3] stays 0 in a 96-row CTA)
    a.lpart[blockIdx.x * 4 + 1] = 0.0f; a.lpart[blockIdx.x * 4 + 2] = 0.0f; a.lpart[blockIdx.x * 4 + 3] = 0.0f;
  }
  if (warp == 0) { tc::tmem_dealloc(tmem, NCOLS_ACC); tc::tmem_dealloc(tmem_a, NCOLS_A); }
}

// ---- weight operand images ------------------------------------------------------------------------------------------------
// One small launch per pass turns theta (flat fp32) into the operand images the two sweeps want - forward: TF32 hi / lo B
// operands in r form + W3 (fwd::OFF_RED floats, the head of the forward kernel's shared memory); adjoint: bf16 hi / lo B
// operands stacked along N + W3 (bwd::WIMG_BYTES, the region [W_BASE, OFF_W3 + 24) of the adjoint kernel's shared memory).  Every
// CTA of the sweeps then fetches its image with ONE TMA bulk copy (cp.async.bulk + mbarrier) instead of ~2000 scalar loads,
// splits and scattered shared-memory stores.
template <int ACT>
__global__ void __launch_bounds__(256) reg_stage_operands_kernel(const PricingArgs a, float* __restrict__ img_f, float* __restrict__ img_b) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int H = a.netA.H, nin = a.netA.nin;
  const float* __restrict__ th = a.theta + a.netA.ext_off;
  const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H;
  if (blockIdx.x == 0) {
    using namespace fwd;
    float* const smem = img_f;
    float* const w3s = smem + OFF_W3;
    constexpr float CS = ACT == ACT_TANH ? 2.885390081777927f : 1.0f;        // pre-activation scale 2 log2(e)
    constexpr float W2S = ACT == ACT_TANH ? -2.0f * CS : 1.0f, W3S = ACT == ACT_TANH ? -2.0f : 1.0f;
    for (int i = tid; i < OFF_RED; i += nthr) smem[i] = 0.0f;
    __syncthreads();
    // tanh layers run in "r form": the MMA delivers x' = 2 log2(e) x (scale folded into the weights), the thread forms
    // r = 1 / (2^x' + 1) (two MUFU + one add) and hands r - not h = 1 - 2 r - to the next layer, whose weights carry the
    // factor -2 and whose bias row carries b + sum_k W[k][.]; the constant-1 unit is r = 1 exactly (2^-200 flushes to 0).
    for (int e = tid; e < n5; e += nthr) {
      float hi, lo;
      if (e < n1) {                                     // W1[i][j], i >= 1 (the time row lives in the effective bias)
        const int i = e / H, j = e % H;
        if (i >= 1) {
          tc::split_tf32(CS * th[e], hi, lo);
          smem[W1B_HI + ((i >> 2) * NB + j) * 4 + (i & 3)] = hi;
          smem[W1B_LO + ((i >> 2) * NB + j) * 4 + (i & 3)] = lo;
        }
      } else if (e < n2) {
      } else if (e < n3) {                              // W2[k][j]
        const int k = (e - n2) / H, j = (e - n2) % H;
        tc::split_tf32(W2S * th[e], hi, lo);
        smem[W2B_HI + ((k >> 2) * NB + j) * 4 + (k & 3)] = hi;
        smem[W2B_LO + ((k >> 2) * NB + j) * 4 + (k & 3)] = lo;
      } else if (e < n4) {
      } else {
        w3s[e - n4] = W3S * th[e];                      // W3[k], k < H
      }
    }
    if (tid < H) {                                      // bias row of layer 2 (k = H)
      float b = th[n3 + tid];
      if (ACT == ACT_TANH) for (int k = 0; k < H; ++k) b += th[n2 + k * H + tid];
      float hi, lo;
      tc::split_tf32(CS * b, hi, lo);
      smem[W2B_HI + ((H >> 2) * NB + tid) * 4 + (H & 3)] = hi;
      smem[W2B_LO + ((H >> 2) * NB + tid) * 4 + (H & 3)] = lo;
    }
    if (tid == nthr - 1) {                              // output bias at index 24
      float b = th[n5];
      if (ACT == ACT_TANH) for (int k = 0; k < H; ++k) b += th[n4 + k];
      w3s[24] = b;
    }
  } else {
    using namespace bwd;
    uint4* const u4 = reinterpret_cast<uint4*>(img_b) - W_BASE;         // the image starts at the adjoint's W_BASE
    float* const w3s = img_b + (OFF_W3 - W_BASE * 4);
    for (int i = tid; i < (int)(WIMG_BYTES / 4); i += nthr) img_b[i] = 0.0f;
    __syncthreads();
    unsigned short* const w1 = reinterpret_cast<unsigned short*>(u4 + W1B);
    unsigned short* const w2 = reinterpret_cast<unsigned short*>(u4 + W2B);
    unsigned short* const wt = reinterpret_cast<unsigned short*>(u4 + WTB);
    unsigned short* const w1t = reinterpret_cast<unsigned short*>(u4 + W1T);
    // element (n, k) of a stacked B operand with NH n-rows per copy: hi at n, lo at NH + n
    auto put = [](unsigned short* w, int NH, int n, int k, uint32_t hi, uint32_t lo) {
      w[((k >> 3) * 2 * NH + n) * 8 + (k & 7)] = (unsigned short)hi;
      w[((k >> 3) * 2 * NH + NH + n) * 8 + (k & 7)] = (unsigned short)lo;
    };
    // tanh: r form (bwd::hid_r) - W1 and the layer-2 bias row carry 2 log2(e), W2 carries -4 log2(e), the bias row adds sum_k W2[k][.],
    // W2^T and W3 (which only meet 1 - h^2 = 4 (r - r^2)) carry the 4
    constexpr float CS = ACT == ACT_TANH ? 2.885390081777927f : 1.0f, W2S = ACT == ACT_TANH ? -2.0f * CS : 1.0f,
                    DS = ACT == ACT_TANH ? 4.0f : 1.0f;
    const float one_in = ACT == ACT_TANH ? -200.0f : 1.0f;   // hid_r(one_in) == 1 exactly: the constant-1 unit of H1 / H2
    for (int e = tid; e < n5 + 2; e += nthr) {
      uint32_t hi, lo;
      if (e < n2) {                                   // W1[i][j]: layer-1 B operand [n = j][k = i]; the time row (i = 0) and
        const int i = e < n1 ? e / H : nin, j = e < n1 ? e % H : e - n1;   // b1 (i = nin) live in the per-step effective bias
        tc::split_bf16(th[e], hi, lo);
        if (i < nin) put(w1t, 16, i, j, hi, lo);      // input-gradient B operand [n = i][k = j]
        tc::split_bf16(CS * th[e], hi, lo);
        if (i >= 1 && i < nin) put(w1, NB, j, i, hi, lo);
      } else if (e < n4) {                            // W2[k][j], b2[j] (k = H): layer-2 B operand [n = j][k]
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        if (k < H) {
          tc::split_bf16(DS * th[e], hi, lo);
          put(wt, NB, k, j, hi, lo);                  // W2^T: B operand [n = k][k' = j]
          tc::split_bf16(W2S * th[e], hi, lo);
        } else {                                      // bias row: b2[j] (+ sum_k W2[k][j] in r form)
          float b = th[e];
          if (ACT == ACT_TANH) for (int kk = 0; kk < H; ++kk) b += th[n2 + kk * H + j];
          tc::split_bf16(CS * b, hi, lo);
        }
        put(w2, NB, j, k, hi, lo);
      } else if (e < n5) {
        w3s[e - n4] = DS * th[e];                     // W3[k][0], k < H  (entries >= H stay 0: no delta for the constant unit)
      } else if (e == n5) {                           // (the constant-1 unit of H1 is written with the effective bias)
      } else {                                        // the constant-1 unit of H2: act(one_in * 1) == 1
        tc::split_bf16(one_in, hi, lo);
        put(w2, NB, H, H, hi, lo);
      }
    }
  }
}

// Tile-major record -> the plane layout of fbsdej_solver_loss' trajectory output: X [N+1][D][B].
template <int D>
__global__ void untile_traj_kernel(const float* __restrict__ rec, const float* __restrict__ recN, TileMap map, int B, int N,
                                   float* __restrict__ out) {
  using RL = RecLayout<D>;
  const size_t total = (size_t)(N + 1) * D * map.ntiles * TR;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(t % TR), tile = (int)((t / TR) % map.ntiles), k = (int)((t / ((size_t)TR * map.ntiles)) % D),
              i = (int)(t / ((size_t)TR * map.ntiles * D));
    const int p = tile_base(map, tile) + row;
    if (row >= tile_height(map, tile) || p >= B) continue;
    out[((size_t)i * D + k) * B + p] =
        i < N ? rec[(((size_t)tile * N + i) * RL::NP + RL::P_X + k) * TR + row] : recN[((size_t)tile * RL::NPT + k) * TR + row];
  }
}

}  // namespace rtc

size_t reg_tc_wimg_fwd_floats() { return (size_t)rtc::fwd::OFF_RED; }
size_t reg_tc_wimg_floats() { return (size_t)rtc::fwd::OFF_RED + rtc::bwd::WIMG_BYTES / 4; }
// builds both operand images (forward image first, adjoint image behind it) from a.theta
int launch_reg_stage_operands(const PricingArgs& a, float* img, cudaStream_t st) {
  float* img_b = img + rtc::fwd::OFF_RED;
  if (a.netA.act == ACT_TANH) rtc::reg_stage_operands_kernel<ACT_TANH><<<2, 256, 0, st>>>(a, img, img_b);
  else rtc::reg_stage_operands_kernel<ACT_RELU><<<2, 256, 0, st>>>(a, img, img_b);
  FB_CUDA(cudaGetLastError());
  return 0;
}
size_t reg_tc_backward_smem() { return sizeof(float) * (size_t)rtc::bwd::SMEM_FLOATS; }
size_t reg_tc_forward_smem(int D) { return sizeof(float) * (size_t)(D == 10 ? rtc::fwd::smem_floats<10>() : rtc::fwd::smem_floats<1>()); }

template <class Model, int ACT, bool RNG>
static int launch_fwd_one(const PricingArgs& a, int grid, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)rtc::fwd::smem_floats<Model::D>();
  auto kern = rtc::reg_forward_tc<Model, ACT, RNG>;
  FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 2 * kThreads, smem, st>>>(a);
  FB_CUDA(cudaGetLastError());
  return 0;
}
template <class Model>
static int launch_fwd(const PricingArgs& a, int grid, cudaStream_t st) {
  if constexpr (Model::kBrownian) {
    if (a.rng)
      return a.netA.act == ACT_TANH ? launch_fwd_one<Model, ACT_TANH, true>(a, grid, st) : launch_fwd_one<Model, ACT_RELU, true>(a, grid, st);
  }
  if (a.rng) { set_error("tcgen05 forward: in-kernel increments exist for the Merton model only"); return -1; }
  return a.netA.act == ACT_TANH ? launch_fwd_one<Model, ACT_TANH, false>(a, grid, st) : launch_fwd_one<Model, ACT_RELU, false>(a, grid, st);
}

int launch_reg_tc_forward(int model, int D, const PricingArgs& a, int grid, cudaStream_t st) {
  if (a.netA.H > 22 || a.netA.nout != 1) { set_error("tcgen05 forward: needs H <= 22 and a single network output"); return -1; }
  if (model == 0 && D == 1) return launch_fwd<MertonModel<1>>(a, grid, st);
  if (model == 0 && D == 10) return launch_fwd<MertonModel<10>>(a, grid, st);
  if (model == 1 && D == 1) return launch_fwd<VGModel>(a, grid, st);
  set_error("tcgen05 forward: unsupported (model, d)");
  return -1;
}

template <class Model>
static int launch_bwd(const PricingArgs& a, int grid, cudaStream_t st) {
  const size_t smem = reg_tc_backward_smem();
  if (a.netA.act == ACT_TANH) {
    auto kern = rtc::reg_backward_tc<Model, ACT_TANH>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(a);
  } else {
    auto kern = rtc::reg_backward_tc<Model, ACT_RELU>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(a);
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

int launch_reg_tc_backward(int model, int D, const PricingArgs& a, int grid, cudaStream_t st) {
  if (a.netA.H > 22 || a.netA.nout != 1) { set_error("tcgen05 adjoint: needs H <= 22 and a single network output"); return -1; }
  if (model == 0 && D == 1) return launch_bwd<MertonModel<1>>(a, grid, st);
  if (model == 0 && D == 10) return launch_bwd<MertonModel<10>>(a, grid, st);
  if (model == 1 && D == 1) return launch_bwd<VGModel>(a, grid, st);
  set_error("tcgen05 adjoint: unsupported (model, d)");
  return -1;
}

int launch_untile_traj(int D, const float* rec, const float* recN, TileMap map, int B, int N, float* out, cudaStream_t st) {
  const int grid = 148 * 4;
  if (D == 1) rtc::untile_traj_kernel<1><<<grid, 256, 0, st>>>(rec, recN, map, B, N, out);
  else if (D == 10) rtc::untile_traj_kernel<10><<<grid, 256, 0, st>>>(rec, recN, map, B, N, out);
  else { set_error("untile: unsupported d"); return -1; }
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fbsdej

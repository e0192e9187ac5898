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
constexpr uint32_t C_ACC = 0, C_W1 = 48, C_W2 = 80, NCOLS = 128;   // ACC 48 | WG1 32 | WG2 48 columns
constexpr int COL_DOUT = 23;                  // spare column of D2 that carries dL/dy through WG2 (needs H <= 22)
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
  uint32_t* const tslot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 4);
  const int row = threadIdx.x, warp = row >> 5;
  const bool issuer = (row & 31) == 0;
  const int H = a.netA.H, nin = a.netA.nin;

  // ---- one-time set-up: zero the tiles, stage the weights as bf16 hi/lo B operands, TMEM, barriers ---------------
  for (int i = row; i < SMEM_FLOATS; i += kThreads) smem[i] = 0.0f;
  __syncthreads();
  {
    const float* __restrict__ th = a.theta + a.netA.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H;
    unsigned short* const w1 = reinterpret_cast<unsigned short*>(u4 + W1B);
    unsigned short* const w2 = reinterpret_cast<unsigned short*>(u4 + W2B);
    unsigned short* const wt = reinterpret_cast<unsigned short*>(u4 + WTB);
    unsigned short* const w1t = reinterpret_cast<unsigned short*>(u4 + W1T);
    // element (n, k) of a stacked B operand with NH n-rows per copy: hi at n, lo at NH + n
    auto put = [](unsigned short* w, int NH, int n, int k, uint32_t hi, uint32_t lo) {
      w[((k >> 3) * 2 * NH + n) * 8 + (k & 7)] = (unsigned short)hi;
      w[((k >> 3) * 2 * NH + NH + n) * 8 + (k & 7)] = (unsigned short)lo;
    };
    const float one_in = ACT == ACT_TANH ? 20.0f : 1.0f;   // act(one_in) == 1 exactly: the constant-1 unit of H1 / H2
    for (int e = row; e < n5 + 2; e += kThreads) {
      uint32_t hi, lo;
      if (e < n2) {                                   // W1[i][j]: layer-1 B operand [n = j][k = i]; the time row (i = 0) and
        const int i = e < n1 ? e / H : nin, j = e < n1 ? e % H : e - n1;   // b1 (i = nin) live in the per-step effective bias
        tc::split_bf16(th[e], hi, lo);
        if (i >= 1 && i < nin) put(w1, NB, j, i, hi, lo);
        if (i < nin) put(w1t, 16, i, j, hi, lo);      // input-gradient B operand [n = i][k = j]
      } else if (e < n4) {                            // W2[k][j], b2[j] (k = H): layer-2 B operand [n = j][k]
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        tc::split_bf16(th[e], hi, lo);
        put(w2, NB, j, k, hi, lo);
        if (k < H) put(wt, NB, k, j, hi, lo);         // W2^T: B operand [n = k][k' = j]
      } else if (e < n5) {
        w3s[e - n4] = th[e];                          // W3[k][0], k < H  (entries >= H stay 0: no delta for the constant unit)
      } else if (e == n5) {                           // (the constant-1 unit of H1 is written with the effective bias)
      } else {                                        // the constant-1 unit of H2: act(one_in * 1) == 1
        tc::split_bf16(one_in, hi, lo);
        put(w2, NB, H, H, hi, lo);
      }
    }
  }
  if (warp == 0) tc::tmem_alloc(tslot, NCOLS);
  if (row == 0) { tc::mbar_init(bar_f, 1); tc::mbar_init(bar_w, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16);
  // thread j <= H owns the effective layer-1 bias of hidden unit j: c_j(t) = t W1[0][j] + b1[j] (fp32, then split)
  float w0 = 0.0f, b1v = 0.0f;
  if (row < H) { w0 = a.theta[a.netA.ext_off + row]; b1v = a.theta[a.netA.ext_off + nin * H + row]; }
  const int bias_idx = ((nin >> 3) * 2 * NB + row) * 8 + (nin & 7);   // hi copy; the lo copy is NB n-rows further
  const uint32_t sbase = tc::smem_u32(u4);
  auto sa = [&](int off_u4) { return sbase + (uint32_t)off_u4 * 16u; };
  uint32_t phase_f = 0, phase_w = 0, pending_w = 0, started = 0;
  auto wait_f = [&]() { tc::mbar_wait(bar_f, phase_f); phase_f ^= 1; tc::tc_fence_after(); };

  const float invB = a.inv_B, invBN = a.inv_B / (float)a.N, rdt = a.r * a.dt;
  const int ntiles = (a.B + TR - 1) / TR;
  const uint32_t step_bytes = (uint32_t)(RL::NP * TR * sizeof(float));
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float msk = (tile * TR + row < a.B) ? 1.0f : 0.0f;
    const float* const rec0 = a.rec + (size_t)tile * a.N * RL::NP * TR + row;
    const float* const recN = a.recN + (size_t)tile * RL::NPT * TR + row;
    float Xbar[D];
    float Esum = 0.0f, rb_next = 0.0f;                  // MultiStep: running sum of e_k;  SumLocal: 2 rho_i / B of the step above
    {
      float X[D];
#pragma unroll
      for (int k = 0; k < D; ++k) X[k] = recN[k * TR];
      float gbar;
      if (a.scheme == SCH_MULTISTEP) {
        Esum = recN[D * TR];                            // sum_k e_k ; d loss / d g = -2/(NB) sum_k e_k
        gbar = -2.0f * Esum * invBN;
      } else {
        rb_next = 2.0f * rec0[((size_t)(a.N - 1) * RL::NP + RL::P_SCH) * TR] * invB;
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
    float Xn[D], En[D], s_n, dA_n, sch_n = 0.0f;
    auto load_step = [&](int i) {
      const float* const rs = rec0 + (size_t)i * RL::NP * TR;
#pragma unroll
      for (int k = 0; k < D; ++k) { Xn[k] = rs[(RL::P_X + k) * TR]; En[k] = rs[(RL::P_E + k) * TR]; }
      s_n = rs[RL::P_S * TR]; dA_n = rs[RL::P_DA * TR];
      // MultiStep: e_i ; SumLocal: rho_{i-1} (the record of the step below; for i = 0 the value is unused)
      sch_n = rs[(a.scheme == SCH_MULTISTEP || i == 0) ? RL::P_SCH * TR : (RL::P_SCH - RL::NP) * TR];
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
          tc::split_bf16(row < H ? fmaf(tf, w0, b1v) : (ACT == ACT_TANH ? 20.0f : 1.0f), hi, lo);
          reinterpret_cast<unsigned short*>(u4 + W1B)[bias_idx] = (unsigned short)hi;
          reinterpret_cast<unsigned short*>(u4 + W1B)[bias_idx + NB * 8] = (unsigned short)lo;
        }
      }
      publish();
      if (warp == 0 && issuer) {
        tc::tc_fence_after();
        gemm_k<1, NB>(tmem + C_ACC, sa(XA_HI), sa(XA_LO), sa(W1B));
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- h1 -> L2 ---------------------------------------------------------------------------------------------
      float h1[24];
      load_acc<NB, 24>(lane_base + C_ACC, h1);
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) h1[8 * c8 + q] = actf<ACT>(h1[8 * c8 + q]);
        tc::store_bf16x8(u4 + H1_HI, u4 + H1_LO, c8, row, h1 + 8 * c8);
      }
      publish();
      if (warp == 1 && issuer) {
        tc::tc_fence_after();
        gemm_k<2, NB>(tmem + C_ACC, sa(H1_HI), sa(H1_LO), sa(W2B));
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
          const float h = actf<ACT>(t8[q]);
          t8[q] = h;
          d2[q] = dout * w8[q] * dactf<ACT>(h);
        }
        if (c8 == 2) d2[COL_DOUT - 16] = dout;
        tc::store_bf16x8(u4 + H2_HI, u4 + H2_LO, c8, row, t8);
        tc::store_bf16x8(u4 + D2_HI, u4 + D2_LO, c8, row, d2);
      }
      publish();
      if (warp == 2 && issuer) {
        tc::tc_fence_after();
        gemm_rows_stacked<48>(tmem + C_W2, sa(H1_HI), sa(D2_HI), started ? 1u : 0u);   // [H1 | H2 (hi) | H1 | H2 (lo)]^T [D2 hi | lo]
        gemm_k<2, NB>(tmem + C_ACC, sa(D2_HI), sa(D2_LO), sa(WTB));
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- delta 1 -> DX, WG1 ---------------------------------------------------------------------------------
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float t8[8], q8[8];
        tc::tmem_ld8(lane_base + C_ACC + 8 * c8, t8);
        tc::tmem_ld8(lane_base + C_ACC + NB + 8 * c8, q8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) t8[q] = (t8[q] + q8[q]) * dactf<ACT>(h1[8 * c8 + q]);
        tc::store_bf16x8(u4 + D1_HI, u4 + D1_LO, c8, row, t8);
      }
      publish();
      if (warp == 3 && issuer) {
        tc::tc_fence_after();
        gemm_k<2, 16>(tmem + C_ACC, sa(D1_HI), sa(D1_LO), sa(W1T));
        tc::mma_commit(bar_f);
        gemm_rows_stacked<32>(tmem + C_W1, sa(D1_HI), sa(XA_HI), started ? 1u : 0u);   // [D1 hi | D1 lo]^T [X hi | X lo] = dW1^T
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
    if (started && pass == 0) {                        // lanes: D1 hi j, D1 lo 24 + j; columns: X hi i, X lo 16 + i
      for (int e = row; e < (nin + 1) * H; e += kThreads) {
        const int i = e / H, j = e % H;                // i = nin: b1
        g[e] = S[j * SW + i] + S[j * SW + 16 + i] + S[(24 + j) * SW + i];
      }
    } else if (started) {                              // lanes: H1 hi 0..23, H2 hi 24..47, H1 lo 48..71, H2 lo 72..95
      for (int e = row; e < (H + 1) * H; e += kThreads) {
        const int k = e / H, j = e % H;                // k = H: b2
        g[o2 + e] = S[k * SW + j] + S[k * SW + 24 + j] + S[(48 + k) * SW + j];
      }
      if (row <= H)                                    // dW3[k] (k = H: b3) rides in column COL_DOUT of D2
        g[o3 + row] = S[(24 + row) * SW + COL_DOUT] + S[(24 + row) * SW + 24 + COL_DOUT] + S[(72 + row) * SW + COL_DOUT];
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
// the tile), so no operand of the split GEMMs is larger than O(1).  The closed-form coupling A(i, X), the exponentials of
// the Euler step and the record stores are issued between an MMA's launch and the wait on its mbarrier.
namespace fwd {
// shared memory: only the B operands (weights) - the activations go to tensor memory
constexpr int W1B_HI = 0, W1B_LO = W1B_HI + 4 * NB * 4, W2B_HI = W1B_LO + 4 * NB * 4, W2B_LO = W2B_HI + 6 * NB * 4,
              OFF_W3 = W2B_LO + 6 * NB * 4 + 32 /* N = 32 reads 8 n-rows past the last chunk */, OFF_RED = OFF_W3 + 32,
              OFF_BAR = OFF_RED + 8, OFF_THR = OFF_BAR + 8 /* 64 Poisson thresholds (fused RNG) */, SMEM_FLOATS = OFF_THR + 64;
static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
// tensor memory: two allocations (32 + 64 = 96 columns, so that five CTAs fit the 512 columns of an SM): the accumulator,
// and the A operand hi (X: 16, H1: 24 columns) | lo
constexpr uint32_t NCOLS_ACC = 32, NCOLS_A = 64;

}  // namespace fwd

// OCC = resident CTAs per SM the kernel is compiled for: 5 (<= 102 registers) pays off when every SM gets at least five
// tiles; with fewer tiles (B = 2^16: 3.5 per SM) the 4-CTA build with its larger register budget is faster.
// RNG: the sweep draws the Merton increments itself (one Philox block per asset pair, in the shadow of the first MMA) - the
// simulation kernel, its 8 d bytes per path-step of stores and this kernel's loads of them disappear from the step.
template <class Model, int ACT, int OCC, bool RNG>
__global__ void __launch_bounds__(kThreads, OCC) reg_forward_tc(const PricingArgs a) {
  constexpr int D = Model::D;
  using RL = RecLayout<D>;
  using namespace fwd;
  static_assert(D + 2 <= 16, "the X tile holds 16 features");
  extern __shared__ __align__(1024) float smem[];
  float* const w3s = smem + OFF_W3;
  float* const red = smem + OFF_RED;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* const tslot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 2);
  const int row = threadIdx.x, warp = row >> 5;
  const bool issuer = (row & 31) == 0;
  const int H = a.netA.H, nin = a.netA.nin;
  const float one_in = ACT == ACT_TANH ? 20.0f : 1.0f;

  for (int i = row; i < SMEM_FLOATS; i += kThreads) smem[i] = 0.0f;
  __syncthreads();
  float w0 = 0.0f, b1v = 0.0f;                          // thread j <= H owns the effective bias of hidden unit j
  {
    const float* __restrict__ th = a.theta + a.netA.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H;
    if (row < H) { w0 = th[row]; b1v = th[n1 + row]; }
    for (int e = row; e <= n5 + 1; e += kThreads) {
      float hi, lo;
      if (e < n1) {                                     // W1[i][j], i >= 1 (the time row lives in the effective bias)
        const int i = e / H, j = e % H;
        if (i >= 1) {
          tc::split_tf32(th[e], hi, lo);
          smem[W1B_HI + ((i >> 2) * NB + j) * 4 + (i & 3)] = hi;
          smem[W1B_LO + ((i >> 2) * NB + j) * 4 + (i & 3)] = lo;
        }
      } else if (e < n2) {
      } else if (e < n4) {                              // W2[k][j], b2[j] (k = H)
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        tc::split_tf32(th[e], hi, lo);
        smem[W2B_HI + ((k >> 2) * NB + j) * 4 + (k & 3)] = hi;
        smem[W2B_LO + ((k >> 2) * NB + j) * 4 + (k & 3)] = lo;
      } else if (e <= n5) {
        w3s[e < n5 ? e - n4 : 24] = th[e];              // W3[k], k < H; b3 at index 24
      } else {
        tc::split_tf32(one_in, hi, lo);
        smem[W2B_HI + ((H >> 2) * NB + H) * 4 + (H & 3)] = hi;
        smem[W2B_LO + ((H >> 2) * NB + H) * 4 + (H & 3)] = lo;
      }
    }
  }
  uint32_t* const sthr = reinterpret_cast<uint32_t*>(smem + OFF_THR);
  if (RNG && row < 64) sthr[row] = row < a.npois ? a.pois_thr[row] : 0xffffffffu;
  if (warp == 0) { tc::tmem_alloc(tslot, NCOLS_ACC, false); tc::tmem_alloc(tslot + 1, NCOLS_A); }
  if (row == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tslot[0], tmem_a = tslot[1];
  const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16), lane_a = tmem_a + ((uint32_t)(row & ~31) << 16);
  const uint32_t sbase = tc::smem_u32(smem);
  auto sa = [&](int off_f) { return sbase + (uint32_t)off_f * 4u; };
  uint32_t phase = 0;
  auto wait_mma = [&]() { tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after(); };
  const int bias_idx = ((nin >> 2) * NB + row) * 4 + (nin & 3);   // W1B[n = row][k = nin]

  const size_t sB = (size_t)a.B;
  const float rdt = a.r * a.dt;
  float lsum = 0.0f;
  const uint32_t rng_iter = RNG ? (a.iter_ptr ? *a.iter_ptr : a.iteration) : 0u;
  const uint32_t t0 = RNG ? sthr[0] : 0u, t1 = RNG ? sthr[1] : 0u;
  const float inv_w1 = t1 > t0 ? 1.0f / (float)(t1 - t0) : 0.0f;
  const int ntiles = (a.B + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * TR + row;
    const bool valid = p0 < a.B;
    const int p = valid ? p0 : a.B - 1;
    const uint32_t gid = a.path_offset + (uint32_t)p;
    float* const rec0 = a.rec + (size_t)tile * a.N * RL::NP * TR + row;
    float X[D];
#pragma unroll
    for (int k = 0; k < D; ++k) X[k] = a.x0;
    float Cpre = 0.0f;                               // MultiStep: sum_{j<i} toAdd_j
    float yprev = 0.0f, aprev = 0.0f, lloc = 0.0f;   // SumLocal
    // the increments of step i are loaded one step ahead (right after the first MMA of step i - 1 is issued)
    float Wn[Model::kBrownian ? D : 1], Jn[D];       // raw values: the combine happens where they are consumed
    auto load_step = [&](int i) {
      if (RNG) return;
      const float* __restrict__ pw = a.dW + (size_t)i * D * sB + p;
      const float* __restrict__ pj = a.J + (size_t)i * D * sB + p;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (Model::kBrownian) Wn[Model::kBrownian ? k : 0] = pw[(size_t)k * sB];
        Jn[k] = pj[(size_t)k * sB];
      }
    };
    load_step(0);
    // RNG: the exponentials e^{drift dt + sig dW + J} of step i + 1 are drawn during step i, two asset pairs in the shadow
    // of the first MMA and the rest in the shadow of the second one
    float En[RNG ? D : 1];
    auto draw = [&](int i, int kp_lo, int kp_hi) {
#pragma unroll
      for (int kp = 0; kp < (D + 1) / 2; ++kp) {
        if (kp < kp_lo || kp >= kp_hi) continue;
        const MertonCell c = merton_cell(gid, ((uint32_t)i << 8) | (uint32_t)kp, rng_iter, STREAM_PATH, a.seed_lo, a.seed_hi, t0, t1,
                                         inv_w1, a.sqdt, a.muJ, a.sigJ, sthr, a.npois);
        En[RNG ? 2 * kp : 0] = __expf(a.drift_dt + a.sig * c.w0 + c.j0);
        if (2 * kp + 1 < D) En[RNG ? 2 * kp + 1 : 0] = __expf(a.drift_dt + a.sig * c.w1 + c.j1);
      }
    };
    constexpr int KP = (D + 1) / 2, KP1 = KP < 2 ? KP : 2;
    if constexpr (RNG) draw(0, 0, KP);
    for (int i = 0; i < a.N; ++i) {
      const float tf = (a.scheme == SCH_SUMLOCAL && a.stale_time) ? (float)(i == 0 ? 0 : i - 1) : (float)i;
      float* const rs = rec0 + (size_t)i * RL::NP * TR;
      float E[D];
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
        // the TMEM stores go last, right before the wait::st of publish_tmem(): tcgen05.st reads its source registers
        // asynchronously, so any instruction that reuses one of them would stall until the store has drained
        store_tf32x8(lane_a, 0, xin);
        store_tf32x8(lane_a, 1, xin + 8);
      }
      publish_tmem();
      if (warp == 0 && issuer) {
        tc::tc_fence_after();
        gemm_k_tf32<2>(tmem, tmem_a, sa(W1B_HI), sa(W1B_LO));
        tc::mma_commit(bar);
      }
      // ---- independent of the network: closed-form coupling, exponentials, record stores -------------------------
      typename Model::AEval ae;
      Model::eval_A_begin(a, i, X, ae);                  // table loads in flight ...
      if constexpr (RNG) {                               // ... while the stores issue and next step's increments are drawn
#pragma unroll
        for (int k = 0; k < D; ++k) {
          rs[(RL::P_X + k) * TR] = X[k];
          E[k] = En[RNG ? k : 0];
        }
        if (i + 1 < a.N) draw(i + 1, 0, KP1);
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) {                    // ... while the exponentials and the stores issue
          rs[(RL::P_X + k) * TR] = X[k];
          E[k] = __expf(a.drift_dt + (Model::kBrownian ? a.sig * Wn[Model::kBrownian ? k : 0] : 0.0f) + Jn[k]);
        }
      }
      float Ai, dAb;
      Model::eval_A_finish(a, i, ae, Ai, dAb);
      rs[RL::P_DA * TR] = dAb;
      wait_mma();
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float t8[8];
        tc::tmem_ld8(lane_base + 8 * c8, t8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) t8[q] = actf<ACT>(t8[q]);
        store_tf32x8(lane_a, c8, t8);                    // (L1 has completed: the X columns are free)
      }
      publish_tmem();
      if (warp == 1 && issuer) {
        tc::tc_fence_after();
        gemm_k_tf32<3>(tmem, tmem_a, sa(W2B_HI), sa(W2B_LO));
        tc::mma_commit(bar);
      }
#pragma unroll
      for (int k = 0; k < D; ++k) rs[(RL::P_E + k) * TR] = E[k];
      if (i + 1 < a.N) {
        if constexpr (RNG) draw(i + 1, KP1, KP); else load_step(i + 1);
      }
      wait_mma();
      float y_net = w3s[24];
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float t8[8];
        tc::tmem_ld8(lane_base + 8 * c8, t8);
        tc::tmem_ld_wait();
        const float4 wa = ld4(w3s + 8 * c8), wb = ld4(w3s + 8 * c8 + 4);
        const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) y_net = fmaf(actf<ACT>(t8[q]), w8[q], y_net);
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
  const float tot = block_sum(lsum, red);
  if (row == 0) {
    a.lpart[blockIdx.x * 4] = tot;
    a.lpart[blockIdx.x * 4 + 1] = 0.0f; a.lpart[blockIdx.x * 4 + 2] = 0.0f; a.lpart[blockIdx.x * 4 + 3] = 0.0f;
  }
  if (warp == 0) { tc::tmem_dealloc(tmem, NCOLS_ACC); tc::tmem_dealloc(tmem_a, NCOLS_A); }
}

// Tile-major record -> the plane layout of fbsdej_solver_loss' trajectory output: X [N+1][D][B].
template <int D>
__global__ void untile_traj_kernel(const float* __restrict__ rec, const float* __restrict__ recN, int B, int N, float* __restrict__ out) {
  using RL = RecLayout<D>;
  const size_t total = (size_t)(N + 1) * D * B;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % B), k = (int)((t / B) % D), i = (int)(t / ((size_t)B * D));
    const int tile = p / TR, row = p % TR;
    out[t] = i < N ? rec[(((size_t)tile * N + i) * RL::NP + RL::P_X + k) * TR + row] : recN[((size_t)tile * RL::NPT + k) * TR + row];
  }
}

}  // namespace rtc

size_t reg_tc_backward_smem() { return sizeof(float) * (size_t)rtc::bwd::SMEM_FLOATS; }
size_t reg_tc_forward_smem() { return sizeof(float) * (size_t)rtc::fwd::SMEM_FLOATS; }

template <class Model, int ACT, int OCC, bool RNG>
static int launch_fwd_one(const PricingArgs& a, int grid, cudaStream_t st) {
  const size_t smem = reg_tc_forward_smem();
  auto kern = rtc::reg_forward_tc<Model, ACT, OCC, RNG>;
  FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreads, smem, st>>>(a);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int reg_tc_forward_occupancy(int B, int sms) { return (B + TR - 1) / TR >= 5 * sms ? 5 : 4; }
template <class Model>
static int launch_fwd(const PricingArgs& a, int grid, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool five = reg_tc_forward_occupancy(a.B, sms) == 5;
  if constexpr (Model::kBrownian) {
    if (a.rng) {
      if (a.netA.act == ACT_TANH)
        return five ? launch_fwd_one<Model, ACT_TANH, 5, true>(a, grid, st) : launch_fwd_one<Model, ACT_TANH, 4, true>(a, grid, st);
      return five ? launch_fwd_one<Model, ACT_RELU, 5, true>(a, grid, st) : launch_fwd_one<Model, ACT_RELU, 4, true>(a, grid, st);
    }
  }
  if (a.rng) { set_error("tcgen05 forward: in-kernel increments exist for the Merton model only"); return -1; }
  if (a.netA.act == ACT_TANH)
    return five ? launch_fwd_one<Model, ACT_TANH, 5, false>(a, grid, st) : launch_fwd_one<Model, ACT_TANH, 4, false>(a, grid, st);
  return five ? launch_fwd_one<Model, ACT_RELU, 5, false>(a, grid, st) : launch_fwd_one<Model, ACT_RELU, 4, false>(a, grid, st);
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

int launch_untile_traj(int D, const float* rec, const float* recN, int B, int N, float* out, cudaStream_t st) {
  const int grid = 148 * 4;
  if (D == 1) rtc::untile_traj_kernel<1><<<grid, 256, 0, st>>>(rec, recN, B, N, out);
  else if (D == 10) rtc::untile_traj_kernel<10><<<grid, 256, 0, st>>>(rec, recN, B, N, out);
  else { set_error("untile: unsupported d"); return -1; }
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fbsdej

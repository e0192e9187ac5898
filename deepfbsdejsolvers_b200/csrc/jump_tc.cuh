// The jump network on tensor cores inside the jump-scheme kernels (pricing_kernels.cu, template flag JTC): the rows of one
// path-step - the path's own jump and the Monte-Carlo compensator samples (SolversJumpDiff.py:37-41, SolversPureJump.py:34-38)
// - are evaluated 128 at a time as one tcgen05 tile: row r of the tile = sample r, thread r = TMEM lane r.
//
//   JumpTcFwd::eval  : y = W3 . act(W2 . act(W1 x))   both layers 3xTF32, A operands in TMEM (as reg_forward_tc)
//   JumpTcBwd::step  : recompute, delta pass, dL/dx, weight gradients accumulated in TMEM over every tile, step and path of
//                      the CTA, six bf16x3 GEMMs (as reg_backward_tc); read once at kernel end (flush)
//
// Only the first output of the network is evaluated (two-network schemes: the jump network's single output; one-network
// schemes: U of the (U, Z) network at the jumped state); the input row [t, state, jump features, 1] has 16 (d = 1) or 24
// (d = 10) features (template parameter NXC).  The time feature is folded into a per-step effective bias (set_time).
// Building blocks: tc_net.cuh.
#pragma once
#include "tc_net.cuh"

namespace fbsdej {

// NXC: 8-feature chunks of the input row (2: up to 15 inputs + the constant 1, d = 1; 3: up to 23, d = 10)
// Forward evaluations: tanh layers in "r form" (as reg_forward_tc): the GEMM delivers x' = 2 log2(e) x (scale in the staged
// weights), the thread forms r = 1 / (2^x' + 1) and hands r - not h = 1 - 2 r - to the next layer, whose weights carry the factor
// -2 and whose bias b + sum_k W[k][.]; the constant-1 unit is r(-200) = 1.  Two instructions per hidden unit less.  (The adjoint
// block keeps h: r in its tiles would turn the weight gradients into db - 2 sum r d, a difference the jump schemes' accuracy
// cannot afford.)
template <int ACT>
__device__ __forceinline__ float jump_hid_r(float x) {
  if (ACT != ACT_TANH) return fmaxf(x, 0.0f);
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return r;
}

template <int ACT, int NXC>
struct JumpTcFwd {
  static constexpr float CS = ACT == ACT_TANH ? 2.885390081777927f : 1.0f;   // layer inputs: 2 log2(e)
  static constexpr float WS = ACT == ACT_TANH ? -2.0f : 1.0f;                // weights that meet r instead of h
  static constexpr float ONE_IN = ACT == ACT_TANH ? -200.0f : 1.0f;          // jump_hid_r(ONE_IN) == 1 exactly
  static_assert(NXC == 2 || NXC == 3, "input row of 16 or 24 features");
  // shared memory (floats): B operands of the two layers (TF32 hi / lo), W3 + b3, the mbarrier, the TMEM slots
  static constexpr int NBR = rtc::NB, NI = 8 * NXC;
  static constexpr int W1B_HI = 0, W1B_LO = W1B_HI + 2 * NXC * NBR * 4, W2B_HI = W1B_LO + 2 * NXC * NBR * 4, W2B_LO = W2B_HI + 6 * NBR * 4,
                       OFF_W3 = W2B_LO + 6 * NBR * 4 + 32, OFF_BAR = OFF_W3 + 32, FLOATS = OFF_BAR + 8;
  float* sm;
  uint64_t* bar;
  uint32_t tmem, tmem_a, lane_base, lane_a, phase, sbase;
  float w0, b1v;
  int H, nin, bias_idx;

  __device__ void init(float* smem, const float* __restrict__ theta, const NetRt& rt) {   // all threads; ends with a barrier
    sm = smem; H = rt.H; nin = rt.nin; phase = 0;
    bar = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 2);
    const int row = threadIdx.x;
    for (int i = row; i < FLOATS; i += kThreads) sm[i] = 0.0f;
    __syncthreads();
    const float* __restrict__ th = theta + rt.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, nout = rt.nout;
    w0 = 0.0f; b1v = 0.0f;
    if (row < H) { w0 = CS * th[row]; b1v = CS * th[n1 + row]; }
    if (row < H) sm[OFF_W3 + row] = WS * th[n4 + row * nout];                // W3[k][0], k < H
    if (row == H) {                                                          // b3[0] at index 24 (r form: + sum_k W3[k][0])
      float b = th[n4 + H * nout];
      if (ACT == ACT_TANH) for (int k = 0; k < H; ++k) b += th[n4 + k * nout];
      sm[OFF_W3 + 24] = b;
    }
    for (int e = row; e < n4; e += kThreads) {
      float hi, lo;
      if (e < n1) {
        const int i = e / H, j = e % H;
        if (i >= 1) {
          tc::split_tf32(CS * th[e], hi, lo);
          sm[W1B_HI + ((i >> 2) * NBR + j) * 4 + (i & 3)] = hi;
          sm[W1B_LO + ((i >> 2) * NBR + j) * 4 + (i & 3)] = lo;
        }
      } else if (e < n2) {
      } else {
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        float v = CS * WS * th[e];
        if (k == H) {                                                        // bias row (r form: + sum_k W2[k][j])
          v = th[e];
          if (ACT == ACT_TANH) for (int kk = 0; kk < H; ++kk) v += th[n2 + kk * H + j];
          v *= CS;
        }
        tc::split_tf32(v, hi, lo);
        sm[W2B_HI + ((k >> 2) * NBR + j) * 4 + (k & 3)] = hi;
        sm[W2B_LO + ((k >> 2) * NBR + j) * 4 + (k & 3)] = lo;
      }
    }
    if (row < 32) { tc::tmem_alloc(tslot, 32, false); tc::tmem_alloc(tslot + 1, 64); }
    if (row == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tmem = tslot[0]; tmem_a = tslot[1];
    lane_base = tmem + ((uint32_t)(row & ~31) << 16);
    lane_a = tmem_a + ((uint32_t)(row & ~31) << 16);
    sbase = tc::smem_u32(sm);
    bias_idx = ((nin >> 2) * NBR + row) * 4 + (nin & 3);
  }
  // effective layer-1 bias c_j = t W1[0][j] + b1[j]; visible to the MMA after the next eval's first barrier
  __device__ __forceinline__ void set_time(float t) {
    const int row = threadIdx.x;
    if (row <= H) {
      float hi, lo;
      tc::split_tf32(row < H ? fmaf(t, w0, b1v) : ONE_IN, hi, lo);
      sm[W1B_HI + bias_idx] = hi;
      sm[W1B_LO + bias_idx] = lo;
    }
  }
  // xin: this row's inputs, xin[0] = 0 (time), xin[nin] = 1.  Every thread of the CTA calls (barriers inside).
  __device__ __forceinline__ float eval(const float (&xin)[NI]) {
    using namespace rtc;
    const int row = threadIdx.x, warp = row >> 5;
#pragma unroll
    for (int c8 = 0; c8 < NXC; ++c8) fwd::store_tf32x8(lane_a, c8, xin + 8 * c8);
    fwd::publish_tmem();
    if (warp == 0 && tc::elect_one()) {
      tc::tc_fence_after();
      fwd::gemm_k_tf32<NXC>(tmem, tmem_a, sbase + W1B_HI * 4, sbase + W1B_LO * 4);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after();
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float t8[8];
      tc::tmem_ld8(lane_base + 8 * c8, t8);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) t8[q] = jump_hid_r<ACT>(t8[q]);
      fwd::store_tf32x8(lane_a, c8, t8);
    }
    fwd::publish_tmem();
    if (warp == 1 && tc::elect_one()) {
      tc::tc_fence_after();
      fwd::gemm_k_tf32<3>(tmem, tmem_a, sbase + W2B_HI * 4, sbase + W2B_LO * 4);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after();
    float y = sm[OFF_W3 + 24];
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float t8[8];
      tc::tmem_ld8(lane_base + 8 * c8, t8);
      tc::tmem_ld_wait();
      const float4 wa = ld4(sm + OFF_W3 + 8 * c8), wb = ld4(sm + OFF_W3 + 8 * c8 + 4);
      const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) y = fmaf(jump_hid_r<ACT>(t8[q]), w8[q], y);
    }
    tc::tc_fence_before();
    return y;
  }
  // ---- separable first layer (one path per thread, two-network schemes) ------------------------------------------------
  // The input row is (t, state, jump features, 1) and the jump features are the same for every row of a tile (all 128 paths meet
  // the same compensator sample), so the layer-1 pre-activation splits into a row part a_b (ONE GEMM per path-step: preact) and
  // a sample part c_m (24 numbers per sample, computed by the caller, uniform over the tile): h1 = act(a_b + scale_b c_m).
  // An evaluation is then one MMA round trip instead of two and needs no input operand.
  // xin: the row's inputs with the jump-feature slots ZEROED; pre: the layer-1 pre-activations (incl. the constant-1 unit).
  __device__ __forceinline__ void preact(const float (&xin)[NI], float (&pre)[24]) {
    using namespace rtc;
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int c8 = 0; c8 < NXC; ++c8) fwd::store_tf32x8(lane_a, c8, xin + 8 * c8);
    fwd::publish_tmem();
    if (warp == 0 && tc::elect_one()) {
      tc::tc_fence_after();
      fwd::gemm_k_tf32<NXC>(tmem, tmem_a, sbase + W1B_HI * 4, sbase + W1B_LO * 4);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after();
    tc::tmem_ld8(lane_base, reinterpret_cast<float (&)[8]>(pre[0]));
    tc::tmem_ld8(lane_base + 8, reinterpret_cast<float (&)[8]>(pre[8]));
    tc::tmem_ld8(lane_base + 16, reinterpret_cast<float (&)[8]>(pre[16]));
    tc::tmem_ld_wait();
    tc::tc_fence_before();
  }
  // c: the sample part (24 floats in shared memory, the same address for every thread), scale: the row's factor on it
  __device__ __forceinline__ float eval_sep(const float (&pre)[24], const float* __restrict__ c, float scale) {
    using namespace rtc;
    const int row = threadIdx.x, warp = row >> 5;
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      const float4 ca = ld4(c + 8 * c8), cb = ld4(c + 8 * c8 + 4);
      const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
      float t8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t8[q] = jump_hid_r<ACT>(fmaf(CS * scale, cv[q], pre[8 * c8 + q]));
      fwd::store_tf32x8(lane_a, c8, t8);
    }
    fwd::publish_tmem();
    if (warp == 1 && tc::elect_one()) {
      tc::tc_fence_after();
      fwd::gemm_k_tf32<3>(tmem, tmem_a, sbase + W2B_HI * 4, sbase + W2B_LO * 4);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after();
    float y = sm[OFF_W3 + 24];
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float t8[8];
      tc::tmem_ld8(lane_base + 8 * c8, t8);
      tc::tmem_ld_wait();
      const float4 wa = ld4(sm + OFF_W3 + 8 * c8), wb = ld4(sm + OFF_W3 + 8 * c8 + 4);
      const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) y = fmaf(jump_hid_r<ACT>(t8[q]), w8[q], y);
    }
    tc::tc_fence_before();
    return y;
  }
  __device__ void finish() {
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc::tmem_dealloc(tmem, 32); tc::tmem_dealloc(tmem_a, 64); }
  }
};

// Sample parts of the separable first layer for the samples [m0, m0 + kThreads) of step i: thread t fills ctab[t][0..24) with
// c_m[j] = sum_k W1[slot0 + k][j] feature(J_m[k]) (sample index nnz = the de-duplicated zero sample; beyond it zeros).
template <class Model>
__device__ __forceinline__ void jump_sample_parts(const PricingArgs& a, int i, int m0, int nnz, float* __restrict__ ctab) {
  constexpr int D = Model::D;
  const int m = m0 + (int)threadIdx.x, H = a.netB.H;
  const float* __restrict__ W = a.theta + a.netB.ext_off + Model::jump_slot0() * H;    // rows slot0 .. of W1 [in][out]
  float f[Model::kJumpSlots];
#pragma unroll
  for (int k = 0; k < Model::kJumpSlots; ++k)
    f[k] = m <= nnz ? Model::jump_feature(a, m < nnz ? a.JMC[((size_t)i * D + k) * a.Mcap + m] : 0.0f) : 0.0f;
  float* __restrict__ out = ctab + (size_t)threadIdx.x * 24;
  for (int j = 0; j < 24; ++j) {
    float acc = 0.0f;
    if (j < H) {
#pragma unroll
      for (int k = 0; k < Model::kJumpSlots; ++k) acc = fmaf(__ldg(W + k * H + j), f[k], acc);
    }
    out[j] = acc;
  }
}

template <int ACT, int NXC>
struct JumpTcBwd {
  static_assert(NXC == 2 || NXC == 3, "input row of 16 or 24 features");
  static constexpr int CH = 128, NBR = rtc::NB, NI = 8 * NXC, KS1 = NXC == 2 ? 1 : 2, NDX = NXC == 2 ? 8 : 16;
  // operand tiles (uint4 offsets from `tiles`; they may alias memory the caller uses between the steps' tile loops)
  // Order matters: a layer GEMM with K = 32 reads one chunk past a 24-feature operand (the weights' rows there are zero, so
  // the chunk only has to hold finite bf16 values): H1_HI -> H2_HI[0], H1_LO -> H2_LO[0] (zeroed by set_time: the caller may
  // have left anything there), D2_HI -> D2_LO[0], D2_LO -> XA_HI[0], D1_HI -> D1_LO[0], D1_LO -> H2_LO[0]; with 24 input
  // features XA_HI -> XA_LO[0], XA_LO -> a pad chunk (zeroed by set_time).
  static constexpr int H1_HI = 0, H2_HI = 3 * CH, H1_LO = 6 * CH, H2_LO = 9 * CH, D2_HI = 12 * CH, D2_LO = 15 * CH, XA_HI = 18 * CH,
                       XA_LO = XA_HI + NXC * CH, XA_PAD = XA_LO + NXC * CH, D1_HI = H2_HI, D1_LO = H1_LO,
                       TILE_U4 = XA_PAD + (NXC == 3 ? CH : 0), TILE_FLOATS = TILE_U4 * 4;
  // weights block (uint4 offsets from `wts`): the stacked B operands (as reg_backward_tc), then W3, the mbarriers, the TMEM slot
  static constexpr int W1B = 0, W2B = W1B + 2 * KS1 * 2 * NBR, WTB = W2B + 4 * 2 * NBR, W1T = WTB + 4 * 2 * NBR, U4_END = W1T + 4 * 2 * NI;
  static constexpr int OFF_W3 = U4_END * 4, OFF_BAR = OFF_W3 + 24, FLOATS = OFF_BAR + 8;
  // tensor memory: layer accumulator (48) | dW1^T (2 NI) | [dW2 | dW3] (48)
  static constexpr uint32_t C_ACC = 0, C_W1 = 48, C_W2 = 48 + 2 * NI, NCOLS = NXC == 2 ? 128 : 256;
  static constexpr int COL_DOUT = 23, SW = 49;
  float* sm;        // weights block
  uint4* u4;        // operand tiles
  uint4* w4;
  uint64_t* bar_f;
  uint64_t* bar_w;
  uint64_t* bar_g;
  uint32_t tmem, lane_base, sbase, wbase, phase_f, phase_w, phase_g, pending_w, started;
  float w0, b1v;
  int H, nin, nout, bias_idx;

  // wts: FLOATS floats; tiles: TILE_FLOATS floats; both 16-byte aligned
  __device__ void init(float* wts, float* tiles, const float* __restrict__ theta, const NetRt& rt) {
    sm = wts; w4 = reinterpret_cast<uint4*>(wts); u4 = reinterpret_cast<uint4*>(tiles); H = rt.H; nin = rt.nin;
    phase_f = phase_w = phase_g = pending_w = started = 0;
    bar_f = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    bar_w = bar_f + 1;
    bar_g = bar_f + 3;                                  // (slot 2 holds the TMEM base)
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 4);
    const int row = threadIdx.x;
    for (int i = row; i < FLOATS; i += kThreads) sm[i] = 0.0f;
    __syncthreads();
    const float* __restrict__ th = theta + rt.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H;
    nout = rt.nout;
    if (row < H) sm[OFF_W3 + row] = th[n4 + row * nout];    // W3[k][0]
    unsigned short* const w1 = reinterpret_cast<unsigned short*>(w4 + W1B);
    unsigned short* const w2 = reinterpret_cast<unsigned short*>(w4 + W2B);
    unsigned short* const wt = reinterpret_cast<unsigned short*>(w4 + WTB);
    unsigned short* const w1t = reinterpret_cast<unsigned short*>(w4 + W1T);
    auto put = [](unsigned short* w, int NH, int n, int k, uint32_t hi, uint32_t lo) {
      w[((k >> 3) * 2 * NH + n) * 8 + (k & 7)] = (unsigned short)hi;
      w[((k >> 3) * 2 * NH + NH + n) * 8 + (k & 7)] = (unsigned short)lo;
    };
    for (int e = row; e < n4 + 1; e += kThreads) {
      uint32_t hi, lo;
      if (e < n2) {
        const int i = e < n1 ? e / H : nin, j = e < n1 ? e % H : e - n1;
        tc::split_bf16(th[e], hi, lo);
        if (i >= 1 && i < nin) put(w1, NBR, j, i, hi, lo);
        if (i < nin) put(w1t, NI, i, j, hi, lo);
      } else if (e < n4) {
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        tc::split_bf16(th[e], hi, lo);
        put(w2, NBR, j, k, hi, lo);
        if (k < H) put(wt, NBR, k, j, hi, lo);
      } else {
        tc::split_bf16(ACT == ACT_TANH ? 20.0f : 1.0f, hi, lo);
        put(w2, NBR, H, H, hi, lo);
      }
    }
    w0 = 0.0f; b1v = 0.0f;
    if (row < H) { w0 = th[row]; b1v = th[n1 + row]; }
    if (row < 32) tc::tmem_alloc(tslot, NCOLS);
    if (row == 0) { tc::mbar_init(bar_f, 1); tc::mbar_init(bar_w, 1); tc::mbar_init(bar_g, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tmem = *tslot;
    lane_base = tmem + ((uint32_t)(row & ~31) << 16);
    sbase = tc::smem_u32(u4);
    wbase = tc::smem_u32(w4);
    bias_idx = ((nin >> 3) * 2 * NBR + row) * 8 + (nin & 7);
  }
  __device__ __forceinline__ uint32_t sa(int off_u4) const { return sbase + (uint32_t)off_u4 * 16u; }
  __device__ __forceinline__ uint32_t sw(int off_u4) const { return wbase + (uint32_t)off_u4 * 16u; }
  __device__ __forceinline__ void wait_f() { tc::mbar_wait(bar_f, phase_f); phase_f ^= 1; tc::tc_fence_after(); }
  __device__ __forceinline__ void drain_w() {
    if (pending_w) { tc::mbar_wait(bar_w, phase_w); phase_w ^= 1; pending_w = 0; }
  }
  // effective layer-1 bias of the step; the previous tile's GEMMs that read the bias row have completed (their results were
  // waited for), so the row can be rewritten right away; the next step()'s first barrier publishes it
  __device__ __forceinline__ void set_time(float t) {
    const int row = threadIdx.x;
    if (row <= H) {
      uint32_t hi, lo;
      tc::split_bf16(row < H ? fmaf(t, w0, b1v) : (ACT == ACT_TANH ? 20.0f : 1.0f), hi, lo);
      reinterpret_cast<unsigned short*>(w4 + W1B)[bias_idx] = (unsigned short)hi;
      reinterpret_cast<unsigned short*>(w4 + W1B)[bias_idx + NBR * 8] = (unsigned short)lo;
    }
    u4[H2_HI + row] = make_uint4(0u, 0u, 0u, 0u);       // K padding of the first tile's layer-2 GEMM
    u4[H2_LO + row] = make_uint4(0u, 0u, 0u, 0u);
    if (NXC == 3) u4[XA_PAD + row] = make_uint4(0u, 0u, 0u, 0u);
  }
  // One tile: xin = this row's inputs (xin[0] = time for dW1, xin[nin] = 1), dout = adjoint of the row's output (0 for rows
  // that do not count).  dx[i] = dL/d xin[i], i < 8.  Every thread of the CTA calls.
  __device__ __forceinline__ void step(const float (&xin)[NI], float dout, float (&dx)[NDX]) {
    using namespace rtc;
    const int row = threadIdx.x, warp = row >> 5;
    drain_w();                                         // WG1 of the previous tile read X, D1
#pragma unroll
    for (int c8 = 0; c8 < NXC; ++c8) tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, c8, row, xin + 8 * c8);
    publish();
    if (warp == 0 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<KS1, NBR, true>(tmem + C_ACC, sa(XA_HI), sa(XA_LO), sw(W1B));
      tc::mma_commit(bar_f);
    }
    wait_f();
    float h1[24];
    load_acc<NBR, 24>(lane_base + C_ACC, h1);
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) h1[8 * c8 + q] = actf<ACT>(h1[8 * c8 + q]);
      tc::store_bf16x8(u4 + H1_HI, u4 + H1_LO, c8, row, h1 + 8 * c8);
    }
    publish();
    if (warp == 1 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<2, NBR, true>(tmem + C_ACC, sa(H1_HI), sa(H1_LO), sw(W2B));
      tc::mma_commit(bar_f);
    }
    wait_f();
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float t8[8], d2[8];
      {
        float q8[8];
        tc::tmem_ld8(lane_base + C_ACC + 8 * c8, t8);
        tc::tmem_ld8(lane_base + C_ACC + NBR + 8 * c8, q8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) t8[q] += q8[q];
      }
      const float4 wa = ld4(sm + OFF_W3 + 8 * c8), wb = ld4(sm + OFF_W3 + 8 * c8 + 4);
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
    if (warp == 2 && tc::elect_one()) {
      tc::tc_fence_after();
      // the input-gradient GEMM first (it is on the tile's chain; the tensor pipe is in order), the weight-gradient GEMM behind
      // it on its own barrier (D1 overwrites tiles it reads)
      gemm_k<2, NBR, true>(tmem + C_ACC, sa(D2_HI), sa(D2_LO), sw(WTB));
      tc::mma_commit(bar_f);
      gemm_rows_stacked<48>(tmem + C_W2, sa(H1_HI), sa(D2_HI), started ? 1u : 0u);
      tc::mma_commit(bar_g);
    }
    wait_f();
    {
      float d1[24];
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float q8[8];
        tc::tmem_ld8(lane_base + C_ACC + 8 * c8, reinterpret_cast<float (&)[8]>(d1[8 * c8]));
        tc::tmem_ld8(lane_base + C_ACC + NBR + 8 * c8, q8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) d1[8 * c8 + q] = (d1[8 * c8 + q] + q8[q]) * dactf<ACT>(h1[8 * c8 + q]);
      }
      tc::mbar_wait(bar_g, phase_g); phase_g ^= 1;
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) tc::store_bf16x8(u4 + D1_HI, u4 + D1_LO, c8, row, d1 + 8 * c8);
    }
    publish();
    if (warp == 3 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<2, NI, true>(tmem + C_ACC, sa(D1_HI), sa(D1_LO), sw(W1T));
      tc::mma_commit(bar_f);
      gemm_rows_stacked<2 * NI, 64>(tmem + C_W1, sa(D1_HI), sa(XA_HI), started ? 1u : 0u);   // M = 64: 48 rows [D1 hi; D1 lo]
      tc::mma_commit(bar_w);
    }
    started = 1;
    pending_w = 1;
    wait_f();
    load_acc<NI, NDX>(lane_base + C_ACC, dx);
    tc::tc_fence_before();
  }
  // ---- separable first layer (JumpTcFwd: preact / eval_sep), adjoint ------------------------------------------------------
  // preact: the state part of the layer-1 pre-activation of this row, ONE GEMM per path-step (xin: inputs with the jump slots
  // zeroed, xin[0] = time, xin[nin] = 1).
  __device__ __forceinline__ void preact(const float (&xin)[NI], float (&pre)[24]) {
    using namespace rtc;
    const int row = threadIdx.x, warp = row >> 5;
    drain_w();
#pragma unroll
    for (int c8 = 0; c8 < NXC; ++c8) tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, c8, row, xin + 8 * c8);
    publish();
    if (warp == 0 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<KS1, NBR, true>(tmem + C_ACC, sa(XA_HI), sa(XA_LO), sw(W1B));
      tc::mma_commit(bar_f);
    }
    wait_f();
    load_acc<NBR, 24>(lane_base + C_ACC, pre);
    tc::tc_fence_before();
  }
  // One compensator sample: h1 = act(pre + scale c) (no layer-1 GEMM), layer 2, deltas; the layer-1 delta d1 is ADDED to sumd1
  // (its input gradient and the weight gradient of the state rows are taken once per path-step, finish_state) and d1 . c to
  // dscale; the weight gradient of the jump-feature rows, (sum_b d1_b) (x) feature, is this sample's own: WG1 on the X tile
  // xj = the row's jump-only inputs (zeros outside the jump slots).  Two MMA round trips instead of four.
  __device__ __forceinline__ void step_sep(const float (&pre)[24], const float* __restrict__ c, float scale, const float (&xj)[NI], float dout,
                                           float (&sumd1)[24], float& dscale) {
    using namespace rtc;
    const int row = threadIdx.x, warp = row >> 5;
    drain_w();                                         // WG1 of the previous sample read X, D1 (= the H2_hi / H1_lo tiles)
#pragma unroll
    for (int c8 = 0; c8 < NXC; ++c8) tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, c8, row, xj + 8 * c8);
    float h1[24];
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      const float4 ca = ld4(c + 8 * c8), cb = ld4(c + 8 * c8 + 4);
      const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) h1[8 * c8 + q] = actf<ACT>(fmaf(scale, cv[q], pre[8 * c8 + q]));
      tc::store_bf16x8(u4 + H1_HI, u4 + H1_LO, c8, row, h1 + 8 * c8);
    }
    publish();
    if (warp == 1 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<2, NBR, true>(tmem + C_ACC, sa(H1_HI), sa(H1_LO), sw(W2B));
      tc::mma_commit(bar_f);
    }
    wait_f();
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float t8[8], d2[8];
      {
        float q8[8];
        tc::tmem_ld8(lane_base + C_ACC + 8 * c8, t8);
        tc::tmem_ld8(lane_base + C_ACC + NBR + 8 * c8, q8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) t8[q] += q8[q];
      }
      const float4 wa = ld4(sm + OFF_W3 + 8 * c8), wb = ld4(sm + OFF_W3 + 8 * c8 + 4);
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
    if (warp == 2 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<2, NBR, true>(tmem + C_ACC, sa(D2_HI), sa(D2_LO), sw(WTB));
      tc::mma_commit(bar_f);
      gemm_rows_stacked<48>(tmem + C_W2, sa(H1_HI), sa(D2_HI), started ? 1u : 0u);
      tc::mma_commit(bar_g);
    }
    wait_f();
    {
      float d1[24];
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float q8[8];
        tc::tmem_ld8(lane_base + C_ACC + 8 * c8, reinterpret_cast<float (&)[8]>(d1[8 * c8]));
        tc::tmem_ld8(lane_base + C_ACC + NBR + 8 * c8, q8);
        tc::tmem_ld_wait();
        const float4 ca = ld4(c + 8 * c8), cb = ld4(c + 8 * c8 + 4);
        const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float v = (d1[8 * c8 + q] + q8[q]) * dactf<ACT>(h1[8 * c8 + q]);
          d1[8 * c8 + q] = v;
          sumd1[8 * c8 + q] += v;
          dscale = fmaf(v, cv[q], dscale);
        }
      }
      tc::mbar_wait(bar_g, phase_g); phase_g ^= 1;
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) tc::store_bf16x8(u4 + D1_HI, u4 + D1_LO, c8, row, d1 + 8 * c8);
    }
    tc::tc_fence_before();
    publish();
    if (warp == 3 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_rows_stacked<2 * NI, 64>(tmem + C_W1, sa(D1_HI), sa(XA_HI), started ? 1u : 0u);
      tc::mma_commit(bar_w);
    }
    started = 1;
    pending_w = 1;
  }
  // Once per path-step after its samples: the summed layer-1 delta against the state inputs - input gradient dx and the weight
  // gradient of the time / state / bias rows of W1.
  __device__ __forceinline__ void finish_state(const float (&xin)[NI], const float (&sumd1)[24], float (&dx)[NDX]) {
    using namespace rtc;
    const int row = threadIdx.x, warp = row >> 5;
    drain_w();
#pragma unroll
    for (int c8 = 0; c8 < NXC; ++c8) tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, c8, row, xin + 8 * c8);
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) tc::store_bf16x8(u4 + D1_HI, u4 + D1_LO, c8, row, sumd1 + 8 * c8);
    publish();
    if (warp == 3 && tc::elect_one()) {
      tc::tc_fence_after();
      gemm_k<2, NI, true>(tmem + C_ACC, sa(D1_HI), sa(D1_LO), sw(W1T));
      tc::mma_commit(bar_f);
      gemm_rows_stacked<2 * NI, 64>(tmem + C_W1, sa(D1_HI), sa(XA_HI), started ? 1u : 0u);
      tc::mma_commit(bar_w);
    }
    started = 1;
    pending_w = 1;
    wait_f();
    load_acc<NI, NDX>(lane_base + C_ACC, dx);
    tc::tc_fence_before();
  }
  // TMEM weight gradients, added to g[...] (external flat layout of this network; first output column of W3 / b3).  All
  // threads call; the operand tiles are dead and serve as scratch (128 x SW floats).
  __device__ void flush(float* __restrict__ g) {
    const int row = threadIdx.x;
    drain_w();
    tc::tc_fence_after();
    __syncthreads();
    float* const S = reinterpret_cast<float*>(u4);
    const int o2 = nin * H + H, o3 = o2 + H * H + H;
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
      if (started) {
#pragma unroll
        for (int c8 = 0; c8 < 6; ++c8) {
          if (pass == 0 && c8 >= 2 * NI / 8) break;
          float v[8];
          tc::tmem_ld8(lane_base + (pass == 0 ? C_W1 : C_W2) + 8 * c8, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) S[row * SW + 8 * c8 + q] = v[q];
        }
      }
      __syncthreads();
      if (started && pass == 0) {
        for (int e = row; e < (nin + 1) * H; e += kThreads) {
          const int i = e / H, j = e % H;
          const int lh = rtc::lane_of_row_m64(j), ll = rtc::lane_of_row_m64(24 + j);     // dW1^T accumulates with M = 64
          g[e] += (S[lh * SW + i] + S[lh * SW + NI + i]) + (S[ll * SW + i] + S[ll * SW + NI + i]);   // hi.hi + hi.lo + lo.hi + lo.lo
        }
      } else if (started) {
        for (int e = row; e < (H + 1) * H; e += kThreads) {
          const int k = e / H, j = e % H;
          g[o2 + e] += (S[k * SW + j] + S[k * SW + 24 + j]) + (S[(48 + k) * SW + j] + S[(48 + k) * SW + 24 + j]);
        }
        if (row <= H)
          g[o3 + row * nout] += (S[(24 + row) * SW + COL_DOUT] + S[(24 + row) * SW + 24 + COL_DOUT]) +
                                (S[(72 + row) * SW + COL_DOUT] + S[(72 + row) * SW + 24 + COL_DOUT]);
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (row < 32) tc::tmem_dealloc(tmem, NCOLS);
  }
};

}  // namespace fbsdej

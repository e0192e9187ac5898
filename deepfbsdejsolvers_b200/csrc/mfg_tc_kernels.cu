// tcgen05 kernels of the smart-grid mean-field game: the two networks of a step (projected player 4 -> H -> H -> nout,
// individual player 6 -> H -> H -> nout, nout <= 4) on the tensor cores, for all five MFG loss graphs
// (coupledMFG/MFGSolvers.py: Global :24-47, MultiStep :187-224, SumLocal :328-364, SumLocalReg :469-505, MultiStepReg
// :615-651; model MFGModel.py:35-107).
//
// A CTA is 2 x 128 threads = two ROLES on the same tile of 128 paths (thread r of each role = path r = TMEM lane r):
// role 0 evaluates / differentiates the projected player's network, role 1 the individual player's, with the GEMM
// machinery of the compensator-free pricing kernels (tc_net.cuh, reg_tc_kernels.cu): forward both layers 3xTF32 with the A
// operands in tensor memory, adjoint six bf16x3 GEMMs per network and step with the weight-gradient accumulators resident
// in TMEM.  The output layer (nout columns) stays on the CUDA cores; its weight gradient rides in a small extra operand tile
// (dout, 8 columns) stacked behind D2 along N.  Every thread carries the path's scalar state redundantly; network outputs
// (forward) and state adjoints (backward) are exchanged through shared memory.  The reference's batch is ONE tile (B = 128)
// walking 95 serial steps: a step's latency is what counts.
#include "mfg.cuh"
#include "tc_net.cuh"

namespace fbsdej {
namespace mtc {
using namespace rtc;

constexpr int kT = 2 * kThreads;

struct Ctl { float ah, al; };
// calpha_hat / calpha, MFGModel.py:82-89
__device__ __forceinline__ Ctl controls(const MFGArgs& a, int i, float hQ, float Q, float R, float hY, float Y) {
  Ctl c;
  const float ind = (R <= a.thetaR) ? 1.0f : 0.0f;
  const float ce = a.coeffEqui;
  const float kTheta = a.A + (1.0f - a.pi) * ce * a.p1 + a.K + ce * a.f1 * ind;
  const float mq = a.meanhq[i];
  const float atg = a.stochastic ? a.alphaTarget * mq : a.alphaTarget;
  c.ah = -(1.0f / kTheta) * (a.p0 + a.pi * a.p1 * hQ + ((1.0f - a.pi) * ce * a.p1 + a.K) * hQ + hY +
                             (a.f0 + ce * a.f1 * (hQ - mq - atg)) * ind);
  c.al = -(1.0f / (a.A + a.K)) * (a.K * Q + a.p0 + a.pi * a.p1 * hQ + (1.0f - a.pi) * ce * a.p1 * (hQ + c.ah) + Y +
                                  (a.f0 + ce * a.f1 * (hQ - mq + c.ah - atg)) * ind);
  return c;
}
__device__ __forceinline__ float block_sum2(float v, float* red) {   // over the 256 threads, valid in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// ---- forward -------------------------------------------------------------------------------------------------------
namespace f {
// per role (floats): B operands of the two layers (TF32 hi / lo), output layer W3 [24][4] + b3 [4]
constexpr int W1B_HI = 0, W1B_LO = W1B_HI + 4 * NB * 4, W2B_HI = W1B_LO + 4 * NB * 4, W2B_LO = W2B_HI + 6 * NB * 4,
              OFF_W3 = W2B_LO + 6 * NB * 4 + 32, OFF_B3 = OFF_W3 + 96, ROLE_FLOATS = OFF_B3 + 8;
// CTA-wide: block-sum scratch, exchanged network outputs [parity][role][128 rows][4], two mbarriers, the TMEM slot
constexpr int OFF_RED = 2 * ROLE_FLOATS, OFF_OUT = OFF_RED + 8, OFF_BAR = OFF_OUT + 2 * 2 * TR * 4, SMEM_FLOATS = OFF_BAR + 8;
static_assert((OFF_BAR % 2) == 0 && (ROLE_FLOATS % 4) == 0, "alignment");
// tensor memory per role: accumulator 32 | A operand hi 32 | lo 32 (fwd::C_AHI / C_ALO relative to the A base)
constexpr uint32_t ROLE_COLS = 128, C_A = 32, NCOLS = 256;
}  // namespace f

// Forward sweep: tanh layers in "r form" (as reg_tc_kernels.cu): the GEMM delivers x' = 2 log2(e) x (scale in the staged weights),
// the thread forms r = 1 / (2^x' + 1) and hands r - not h = 1 - 2 r - to the next layer, whose weights carry the factor -2 and
// whose bias row b + sum_k W[k][.]; the constant-1 unit is r(-200) = 1.  Two instructions per hidden unit less.  (The adjoint
// sweep keeps h: with r in its tiles the weight gradients become db - 2 sum r d, and that difference costs a factor ~3 in
// accuracy - 5.7e-5 of the largest component against the 5e-5 bound of tests/test_tc_gpu.py.)
template <int ACT>
__device__ __forceinline__ float hid_r(float x) {
  if (ACT != ACT_TANH) return fmaxf(x, 0.0f);
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return r;
}
template <int ACT> struct RForm {
  static constexpr float CS = ACT == ACT_TANH ? 2.885390081777927f : 1.0f;     // layer inputs: 2 log2(e)
  static constexpr float WS = ACT == ACT_TANH ? -2.0f : 1.0f;                  // weights that meet r instead of h
  static constexpr float ONE_IN = ACT == ACT_TANH ? -200.0f : 1.0f;            // hid_r(ONE_IN) == 1 exactly
};

template <int ACT>
__global__ void __launch_bounds__(kT, 1) mfg_forward_tc(const MFGArgs a) {
  using namespace f;
  extern __shared__ __align__(1024) float smem[];
  const int role = threadIdx.x >> 7, row = threadIdx.x & (TR - 1), warp = row >> 5;
  const NetRt& net = role == 0 ? a.netA : a.netB;
  const int H = net.H, nin = net.nin, nout = net.nout;
  float* const rw = smem + role * ROLE_FLOATS;
  float* const red = smem + OFF_RED;
  float* const outx = smem + OFF_OUT;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + role;
  uint32_t* const tslot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 4);
  using RF = RForm<ACT>;
  const float one_in = RF::ONE_IN;

  for (int i = threadIdx.x; i < SMEM_FLOATS; i += kT) smem[i] = 0.0f;
  __syncthreads();
  float w0 = 0.0f, b1v = 0.0f;                       // thread j <= H owns the effective layer-1 bias of hidden unit j
  {
    const float* __restrict__ th = a.theta + net.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H * nout, n6 = n5 + nout;
    if (row < H) { w0 = RF::CS * th[row]; b1v = RF::CS * th[n1 + row]; }
    for (int e = row; e < n6; e += TR) {
      float hi, lo;
      if (e < n1) {                                  // W1[i][j], i >= 1 (the time row lives in the effective bias)
        const int i = e / H, j = e % H;
        if (i >= 1) {
          tc::split_tf32(RF::CS * th[e], hi, lo);
          rw[W1B_HI + ((i >> 2) * NB + j) * 4 + (i & 3)] = hi;
          rw[W1B_LO + ((i >> 2) * NB + j) * 4 + (i & 3)] = lo;
        }
      } else if (e < n2) {
      } else if (e < n4) {                           // W2[k][j], b2[j] (k = H; r form: + sum_k W2[k][j])
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        float v = RF::CS * RF::WS * th[e];
        if (k == H) {
          v = th[e];
          if (ACT == ACT_TANH) for (int kk = 0; kk < H; ++kk) v += th[n2 + kk * H + j];
          v *= RF::CS;
        }
        tc::split_tf32(v, hi, lo);
        rw[W2B_HI + ((k >> 2) * NB + j) * 4 + (k & 3)] = hi;
        rw[W2B_LO + ((k >> 2) * NB + j) * 4 + (k & 3)] = lo;
      } else if (e < n5) {                           // W3[k][o] -> [k][4]
        rw[OFF_W3 + ((e - n4) / nout) * 4 + (e - n4) % nout] = RF::WS * th[e];
      } else {                                       // b3[o] (r form: + sum_k W3[k][o])
        float v = th[e];
        if (ACT == ACT_TANH) for (int kk = 0; kk < H; ++kk) v += th[n4 + kk * nout + (e - n5)];
        rw[OFF_B3 + (e - n5)] = v;
      }
    }
  }
  if (threadIdx.x < 32) tc::tmem_alloc(tslot, NCOLS);
  if (row == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tslot + role * ROLE_COLS, tmem_a = tmem + C_A;
  const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16), lane_a = tmem_a + ((uint32_t)(row & ~31) << 16);
  const uint32_t sbase = tc::smem_u32(rw);
  auto sa = [&](int off_f) { return sbase + (uint32_t)off_f * 4u; };
  uint32_t phase = 0;
  auto wait_mma = [&]() { tc::mbar_wait(bar, phase); phase ^= 1; tc::tc_fence_after(); };
  const int bias_idx = ((nin >> 2) * NB + row) * 4 + (nin & 3);   // W1B[n = row][k = nin]

  const size_t sB = (size_t)a.B;
  const int c0 = a.has_y ? 1 : 0;
  float lh_sum = 0.0f, li_sum = 0.0f;
  const int ntiles = (a.B + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * TR + row;
    const bool valid = p0 < a.B;
    const bool writer = valid && role == 0;
    const int p = valid ? p0 : a.B - 1;
    float hQ = a.q0, Q = a.q0, R = a.R0, hS = a.S0, S = a.S0;
    float hY = 0.0f, Y = 0.0f;
    if (a.scheme == SCH_GLOBAL) { hY = a.theta[a.y0_off]; Y = a.theta[a.y0_off + 1]; }
    float Ch = 0.0f, Ci = 0.0f;                                                       // MultiStep
    float hyp = 0.0f, ahp = 0.0f, yp = 0.0f, aip = 0.0f, llh = 0.0f, lli = 0.0f;      // SumLocal
    for (int i = 0; i < a.N; ++i) {
      const float tm = (float)i * a.dt;
      const float dW0 = a.dW0[(size_t)i * sB + p], dW = a.dW[(size_t)i * sB + p], dN = a.dN[(size_t)i * sB + p];
      {
        float xin[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) xin[k] = 0.0f;
        if (role == 0) { xin[1] = hQ; xin[2] = hS; xin[3] = R; xin[4] = 1.0f; }                       // getProjectedStates
        else { xin[1] = Q; xin[2] = S; xin[3] = hQ; xin[4] = hS; xin[5] = R; xin[6] = 1.0f; }       // getAllStates
        if (row <= H) {
          float hi, lo;
          tc::split_tf32(row < H ? fmaf(tm, w0, b1v) : one_in, hi, lo);
          rw[W1B_HI + bias_idx] = hi;
          rw[W1B_LO + bias_idx] = lo;
        }
        fwd::store_tf32x8(lane_a, 0, xin);
        fwd::store_tf32x8(lane_a, 1, xin + 8);
      }
      fwd::publish_tmem();
      if (warp == 0 && tc::elect_one()) {
        tc::tc_fence_after();
        fwd::gemm_k_tf32<2>(tmem, tmem_a, sa(W1B_HI), sa(W1B_LO));
        tc::mma_commit(bar);
      }
      wait_mma();
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float t8[8];
        tc::tmem_ld8(lane_base + 8 * c8, t8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) t8[q] = hid_r<ACT>(t8[q]);
        fwd::store_tf32x8(lane_a, c8, t8);
      }
      fwd::publish_tmem();
      if (warp == 1 && tc::elect_one()) {
        tc::tc_fence_after();
        fwd::gemm_k_tf32<3>(tmem, tmem_a, sa(W2B_HI), sa(W2B_LO));
        tc::mma_commit(bar);
      }
      wait_mma();
      float4 o = ld4(rw + OFF_B3);
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        float t8[8];
        tc::tmem_ld8(lane_base + 8 * c8, t8);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float h = hid_r<ACT>(t8[q]);
          const float4 w = ld4(rw + OFF_W3 + (8 * c8 + q) * 4);     // rows >= H are zero
          o.x = fmaf(h, w.x, o.x); o.y = fmaf(h, w.y, o.y); o.z = fmaf(h, w.z, o.z); o.w = fmaf(h, w.w, o.w);
        }
      }
      tc::tc_fence_before();
      float* const ox = outx + (i & 1) * (2 * TR * 4);
      st4(ox + (role * TR + row) * 4, o);
      __syncthreads();                                                               // both networks' outputs of step i
      const float4 oh = ld4(ox + row * 4), oi = ld4(ox + (TR + row) * 4);
      const float oh0 = oh.x, oh1 = oh.y, oh2 = oh.z;
      const float o0 = oi.x, o1 = oi.y, o2 = oi.z, o3 = oi.w;
      const float lamdt = (a.stochastic ? a.beta * (expf(a.alpha * hQ) - 1.0f) : a.jumpFactor) * a.dt;
      const float dNc = dN - lamdt;
      float a_h = -a.dt * (hS * a.C), a_i = -a.dt * (S * a.C);
      if (a.has_z) {
        const float hz0 = c0 ? oh1 : oh0, hgam = c0 ? oh2 : oh1;
        const float z0 = c0 ? o1 : o0, gam = c0 ? o2 : o1, z = c0 ? o3 : o2;
        a_h = a_h + hz0 * dW0 + hgam * dNc;                    // MFGSolvers.py:40 / :203
        a_i = a_i + z0 * dW0 + gam * dNc + z * dW;             // :41 / :204
      }
      const float hYsel = (a.scheme == SCH_GLOBAL) ? hY : oh0;
      const float Ysel = (a.scheme == SCH_GLOBAL) ? Y : o0;
      if (writer) {
        float* tx = a.traj + ((size_t)i * 5) * sB + p;
        tx[0] = hQ; tx[sB] = Q; tx[2 * sB] = R; tx[3 * sB] = hS; tx[4 * sB] = S;
        if (a.trajY) { a.trajY[((size_t)i * 2) * sB + p] = hYsel; a.trajY[((size_t)i * 2 + 1) * sB + p] = Ysel; }
      }
      if (a.scheme == SCH_GLOBAL) {
        hY += a_h; Y += a_i;
      } else if (a.scheme == SCH_MULTISTEP) {
        if (writer) {
          a.sch[((size_t)i * 2 + 0) * sB + p] = hYsel - Ch;
          a.sch[((size_t)i * 2 + 1) * sB + p] = Ysel - Ci;
        }
        Ch += a_h; Ci += a_i;
      } else {
        if (i > 0) {
          const float rh = hYsel - hyp - ahp, ri = Ysel - yp - aip;
          llh = fmaf(rh, rh, llh); lli = fmaf(ri, ri, lli);
          if (writer) { a.sch[((size_t)(i - 1) * 2 + 0) * sB + p] = rh; a.sch[((size_t)(i - 1) * 2 + 1) * sB + p] = ri; }
        }
        hyp = hYsel; ahp = a_h; yp = Ysel; aip = a_i;
      }
      // oneStepFrom, MFGModel.py:58-71 (controls use the states of step i)
      const Ctl c = controls(a, i, hQ, Q, R, hYsel, Ysel);
      hS = hS + c.ah * a.dt;
      S = S + c.al * a.dt;
      R = R + a.dt - (dN > 0.0f ? R : 0.0f);
      const float qn = a.qaver[i + 1];
      hQ = hQ + a.coeffOU * (qn - hQ) * a.dt + a.sig0 * dW0;
      Q = Q + a.coeffOU * (qn - Q) * a.dt + a.sig0 * dW0 + a.sig * dW;
    }
    const float gh = a.h1 + a.h2 * hS, gi = a.h1 + a.h2 * S;
    float lh = 0.0f, li = 0.0f;
    if (a.scheme == SCH_GLOBAL) {
      const float eh = hY - gh, ei = Y - gi;
      lh = eh * eh * a.inv_B; li = ei * ei * a.inv_B;
      if (writer) { a.fin[p] = eh; a.fin[sB + p] = ei; }
    } else if (a.scheme == SCH_MULTISTEP) {
      if (writer) {
        const float Dh = Ch - gh, Di = Ci - gi;
        float seh = 0.0f, sei = 0.0f, s2h = 0.0f, s2i = 0.0f;
        for (int k = 0; k < a.N; ++k) {
          const float eh = a.sch[((size_t)k * 2 + 0) * sB + p] + Dh, ei = a.sch[((size_t)k * 2 + 1) * sB + p] + Di;
          a.sch[((size_t)k * 2 + 0) * sB + p] = eh; a.sch[((size_t)k * 2 + 1) * sB + p] = ei;
          seh += eh; sei += ei; s2h = fmaf(eh, eh, s2h); s2i = fmaf(ei, ei, s2i);
        }
        lh = s2h * (a.inv_B / (float)a.N); li = s2i * (a.inv_B / (float)a.N);
        a.fin[p] = seh; a.fin[sB + p] = sei;
      }
    } else {
      const float rh = gh - hyp - ahp, ri = gi - yp - aip;
      llh = fmaf(rh, rh, llh); lli = fmaf(ri, ri, lli);
      lh = llh * a.inv_B; li = lli * a.inv_B;
      if (writer) { a.sch[((size_t)(a.N - 1) * 2 + 0) * sB + p] = rh; a.sch[((size_t)(a.N - 1) * 2 + 1) * sB + p] = ri; }
    }
    if (writer) {
      float* tx = a.traj + ((size_t)a.N * 5) * sB + p;
      tx[0] = hQ; tx[sB] = Q; tx[2 * sB] = R; tx[3 * sB] = hS; tx[4 * sB] = S;
      if (a.trajY) {
        a.trajY[((size_t)a.N * 2) * sB + p] = (a.scheme == SCH_GLOBAL) ? hY : gh;
        a.trajY[((size_t)a.N * 2 + 1) * sB + p] = (a.scheme == SCH_GLOBAL) ? Y : gi;
      }
      lh_sum += lh; li_sum += li;
    }
  }
  tc::tc_fence_before();
  const float th = block_sum2(lh_sum, red);
  const float ti = block_sum2(li_sum, red);
  if (threadIdx.x == 0) {
    a.lpart[blockIdx.x * 4 + 0] = a.w_hat * th + a.w_ind * ti;
    a.lpart[blockIdx.x * 4 + 1] = th;
    a.lpart[blockIdx.x * 4 + 2] = ti;
    a.lpart[blockIdx.x * 4 + 3] = 0.0f;
  }
  if (threadIdx.x < 32) tc::tmem_dealloc(*tslot, NCOLS);
}

// ---- adjoint -------------------------------------------------------------------------------------------------------
namespace b {
constexpr int CH = 128;                       // uint4 per chunk (128 rows x 16 bytes)
// per role (uint4): operand tiles as in reg_tc_kernels.cu plus the dout tile DO (8 columns) behind D2
constexpr int XA_HI = 0, XA_LO = 2 * CH, H1_HI = 4 * CH, H2_HI = 7 * CH, H1_LO = 10 * CH, H2_LO = 13 * CH, D2_HI = 16 * CH,
              D2_LO = 19 * CH, DO_HI = 22 * CH, DO_LO = 23 * CH, D1_HI = H2_HI, D1_LO = H1_LO, W_BASE = 24 * CH;
constexpr int W1B = W_BASE, W2B = W1B + 2 * 2 * NB, WTB = W2B + 4 * 2 * NB, W1T = WTB + 4 * 2 * NB, W3F = W1T + 4 * 2 * 16,
              ROLE_U4 = W3F + 24 + 8;         // W3 [24][4] floats = 24 uint4 (+ pad)
constexpr int ROLE_FLOATS = ROLE_U4 * 4;
// CTA-wide floats: block-sum scratch, exchanged state adjoints [parity][3][128], mbarriers (2 x bar_f, 2 x bar_w), TMEM slot
constexpr int OFF_RED = 2 * ROLE_FLOATS, OFF_DX = OFF_RED + 8, OFF_BAR = OFF_DX + 2 * 3 * TR, SMEM_FLOATS = OFF_BAR + 12;
static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
// tensor memory per role: accumulator 48 | dW1^T 32 | [dW2 | dW3] 64
constexpr uint32_t C_ACC = 0, C_W1 = 48, C_W2 = 80, ROLE_COLS = 160, NCOLS = 512;
constexpr int SW = 65;                        // row stride of the flush scratch
}  // namespace b

template <int ACT>
__global__ void __launch_bounds__(kT, 1) mfg_backward_tc(const MFGArgs a) {
  using namespace b;
  extern __shared__ __align__(1024) float smem[];
  const int role = threadIdx.x >> 7, row = threadIdx.x & (TR - 1), warp = row >> 5;
  const NetRt& net = role == 0 ? a.netA : a.netB;
  const int H = net.H, nin = net.nin, nout = net.nout;
  float* const rf = smem + role * ROLE_FLOATS;
  uint4* const u4 = reinterpret_cast<uint4*>(rf);
  float* const w3 = rf + W3F * 4;                                  // W3[k][o] as [k][4]
  float* const red = smem + OFF_RED;
  float* const dxx = smem + OFF_DX;
  uint64_t* const bar_f = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + role;
  uint64_t* const bar_w = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + 2 + role;
  uint32_t* const tslot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8);

  for (int i = threadIdx.x; i < SMEM_FLOATS; i += kT) smem[i] = 0.0f;
  __syncthreads();
  {
    const float* __restrict__ th = a.theta + net.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H * nout;
    unsigned short* const w1 = reinterpret_cast<unsigned short*>(u4 + W1B);
    unsigned short* const w2 = reinterpret_cast<unsigned short*>(u4 + W2B);
    unsigned short* const wt = reinterpret_cast<unsigned short*>(u4 + WTB);
    unsigned short* const w1t = reinterpret_cast<unsigned short*>(u4 + W1T);
    auto put = [](unsigned short* w, int NH, int n, int k, uint32_t hi, uint32_t lo) {
      w[((k >> 3) * 2 * NH + n) * 8 + (k & 7)] = (unsigned short)hi;
      w[((k >> 3) * 2 * NH + NH + n) * 8 + (k & 7)] = (unsigned short)lo;
    };
    const float one_in = ACT == ACT_TANH ? 20.0f : 1.0f;
    for (int e = row; e < n5 + 1; e += TR) {
      uint32_t hi, lo;
      if (e < n2) {                                   // W1[i][j]: layer-1 B operand [n = j][k = i] (time row and b1: effective bias)
        const int i = e < n1 ? e / H : nin, j = e < n1 ? e % H : e - n1;
        tc::split_bf16(th[e], hi, lo);
        if (i >= 1 && i < nin) put(w1, NB, j, i, hi, lo);
        if (i < nin) put(w1t, 16, i, j, hi, lo);      // input-gradient B operand [n = i][k = j]
      } else if (e < n4) {                            // W2[k][j], b2[j] (k = H)
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        tc::split_bf16(th[e], hi, lo);
        put(w2, NB, j, k, hi, lo);
        if (k < H) put(wt, NB, k, j, hi, lo);
      } else if (e < n5) {                            // W3[k][o], k < H (rows >= H stay 0: no delta for the constant unit)
        w3[((e - n4) / nout) * 4 + (e - n4) % nout] = th[e];
      } else {                                        // the constant-1 unit of H2 (carries b3 through the weight gradient)
        tc::split_bf16(one_in, hi, lo);
        put(w2, NB, H, H, hi, lo);
      }
    }
  }
  if (threadIdx.x < 32) tc::tmem_alloc(tslot, NCOLS);
  if (row == 0) { tc::mbar_init(bar_f, 1); tc::mbar_init(bar_w, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tslot + role * ROLE_COLS;
  const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16);
  float w0 = 0.0f, b1v = 0.0f;
  if (row < H) { w0 = a.theta[net.ext_off + row]; b1v = a.theta[net.ext_off + nin * H + row]; }
  const int bias_idx = ((nin >> 3) * 2 * NB + row) * 8 + (nin & 7);
  const uint32_t sbase = tc::smem_u32(u4);
  auto sa = [&](int off_u4) { return sbase + (uint32_t)off_u4 * 16u; };
  uint32_t phase_f = 0, phase_w = 0, pending_w = 0, started = 0;
  auto wait_f = [&]() { tc::mbar_wait(bar_f, phase_f); phase_f ^= 1; tc::tc_fence_after(); };

  const size_t sB = (size_t)a.B;
  const int c0 = a.has_y ? 1 : 0;
  const float invB = a.inv_B, invBN = a.inv_B / (float)a.N;
  const float wh = a.w_hat, wi = a.w_ind;
  float y0h = 0.0f, y0i = 0.0f;
  const int ntiles = (a.B + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * TR + row;
    const bool valid = p0 < a.B;
    const int p = valid ? p0 : a.B - 1;
    const float msk = valid ? 1.0f : 0.0f;
    float Eh = 0.0f, Ei = 0.0f;
    float ghbar, gibar, hYbar = 0.0f, Ybar = 0.0f;
    if (a.scheme == SCH_GLOBAL) {
      hYbar = 2.0f * a.fin[p] * invB * wh; Ybar = 2.0f * a.fin[sB + p] * invB * wi;
      ghbar = -hYbar; gibar = -Ybar;
    } else if (a.scheme == SCH_MULTISTEP) {
      Eh = a.fin[p]; Ei = a.fin[sB + p];
      ghbar = -2.0f * Eh * invBN * wh;
      gibar = -2.0f * Ei * invBN * wi;
    } else {
      ghbar = 2.0f * a.sch[((size_t)(a.N - 1) * 2 + 0) * sB + p] * invB * wh;
      gibar = 2.0f * a.sch[((size_t)(a.N - 1) * 2 + 1) * sB + p] * invB * wi;
    }
    float hSbar = ghbar * a.h2, Sbar = gibar * a.h2;
    for (int i = a.N - 1; i >= 0; --i) {
      const float tm = (float)i * a.dt;
      const float* tx = a.traj + ((size_t)i * 5) * sB + p;
      const float hQ = tx[0], Q = tx[sB], R = tx[2 * sB], hS = tx[3 * sB], S = tx[4 * sB];
      const float dW0 = a.dW0[(size_t)i * sB + p], dW = a.dW[(size_t)i * sB + p], dN = a.dN[(size_t)i * sB + p];
      const float lamdt = (a.stochastic ? a.beta * (expf(a.alpha * hQ) - 1.0f) : a.jumpFactor) * a.dt;
      const float dNc = dN - lamdt;
      // adjoint of the controlled states: hS' = hS + ah dt, S' = S + al dt
      const float ind = (R <= a.thetaR) ? 1.0f : 0.0f;
      const float ce = a.coeffEqui;
      const float kTheta = a.A + (1.0f - a.pi) * ce * a.p1 + a.K + ce * a.f1 * ind;
      const float albar = Sbar * a.dt;
      const float dal_dah = -(1.0f / (a.A + a.K)) * ((1.0f - a.pi) * ce * a.p1 + ce * a.f1 * ind);
      const float ahbar = hSbar * a.dt + albar * dal_dah;
      const float cYi = albar * (-1.0f / (a.A + a.K));   // adjoint into the Y fed to oneStepFrom
      const float cYh = ahbar * (-1.0f / kTheta);        // adjoint into hY
      float abh, abi, hyb = 0.0f, yb = 0.0f;
      if (a.scheme == SCH_GLOBAL) {
        abh = hYbar; abi = Ybar;                 // hY_{i+1} = hY_i + a_h
        hYbar += cYh; Ybar += cYi;               // OLD hY_i, Y_i feed the controls (MFGSolvers.py:43)
      } else if (a.scheme == SCH_MULTISTEP) {
        const float eh = a.sch[((size_t)i * 2 + 0) * sB + p], ei = a.sch[((size_t)i * 2 + 1) * sB + p];
        abh = 2.0f * Eh * invBN * wh;
        abi = 2.0f * Ei * invBN * wi;
        hyb = 2.0f * eh * invBN * wh + cYh;
        yb = 2.0f * ei * invBN * wi + cYi;
        Eh -= eh; Ei -= ei;
      } else {
        const float rbh = 2.0f * a.sch[((size_t)i * 2 + 0) * sB + p] * invB * wh;
        const float rbi = 2.0f * a.sch[((size_t)i * 2 + 1) * sB + p] * invB * wi;
        float rbhm = 0.0f, rbim = 0.0f;
        if (i > 0) {
          rbhm = 2.0f * a.sch[((size_t)(i - 1) * 2 + 0) * sB + p] * invB * wh;
          rbim = 2.0f * a.sch[((size_t)(i - 1) * 2 + 1) * sB + p] * invB * wi;
        }
        abh = -rbh; abi = -rbi;
        hyb = rbhm - rbh + cYh;
        yb = rbim - rbi + cYi;
      }
      // direct dependence of the increments on the states: a_h = -dt C hS + ..., a = -dt C S + ...
      hSbar += -a.dt * a.C * abh;
      Sbar += -a.dt * a.C * abi;
      // this role's network: inputs and the adjoints of its outputs
      float xin[16], dd[8];
#pragma unroll
      for (int k = 0; k < 16; ++k) xin[k] = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) dd[k] = 0.0f;
      xin[0] = tm;                                     // (layer 1 takes the time through the effective bias; dW1 needs it here)
      if (role == 0) {
        xin[1] = hQ; xin[2] = hS; xin[3] = R; xin[4] = 1.0f;
        if (a.has_y) dd[0] = hyb * msk;
        if (a.has_z) {
          const float z0b = abh * dW0 * msk, gb = abh * dNc * msk;
          if (c0) { dd[1] = z0b; dd[2] = gb; } else { dd[0] = z0b; dd[1] = gb; }
        }
      } else {
        xin[1] = Q; xin[2] = S; xin[3] = hQ; xin[4] = hS; xin[5] = R; xin[6] = 1.0f;
        if (a.has_y) dd[0] = yb * msk;
        if (a.has_z) {
          const float z0b = abi * dW0 * msk, gb = abi * dNc * msk, zb = abi * dW * msk;
          if (c0) { dd[1] = z0b; dd[2] = gb; dd[3] = zb; } else { dd[0] = z0b; dd[1] = gb; dd[2] = zb; }
        }
      }
      // ---- X tile -> L1 ---------------------------------------------------------------------------------------------
      if (pending_w) { tc::mbar_wait(bar_w, phase_w); phase_w ^= 1; pending_w = 0; }   // WG1 of the step above read X, D1
      tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, 0, row, xin);
      tc::store_bf16x8(u4 + XA_HI, u4 + XA_LO, 1, row, xin + 8);
      if (row <= H) {
        uint32_t hi, lo;
        tc::split_bf16(row < H ? fmaf(tm, w0, b1v) : (ACT == ACT_TANH ? 20.0f : 1.0f), hi, lo);
        reinterpret_cast<unsigned short*>(u4 + W1B)[bias_idx] = (unsigned short)hi;
        reinterpret_cast<unsigned short*>(u4 + W1B)[bias_idx + NB * 8] = (unsigned short)lo;
      }
      publish();
      if (warp == 0 && tc::elect_one()) {
        tc::tc_fence_after();
        gemm_k<1, NB>(tmem + C_ACC, sa(XA_HI), sa(XA_LO), sa(W1B));
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- h1 -> L2 -------------------------------------------------------------------------------------------------
      float h1[24];
      load_acc<NB, 24>(lane_base + C_ACC, h1);
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) h1[8 * c8 + q] = actf<ACT>(h1[8 * c8 + q]);
        tc::store_bf16x8(u4 + H1_HI, u4 + H1_LO, c8, row, h1 + 8 * c8);
      }
      publish();
      if (warp == 1 && tc::elect_one()) {
        tc::tc_fence_after();
        gemm_k<2, NB>(tmem + C_ACC, sa(H1_HI), sa(H1_LO), sa(W2B));
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- h2, delta 2 = (W3 dout) .* act'(h2), dout tile -> [dW2 | dW3], D2 W2^T ---------------------------------------
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
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float h = actf<ACT>(t8[q]);
          const float4 w = ld4(w3 + (8 * c8 + q) * 4);
          t8[q] = h;
          d2[q] = (dd[0] * w.x + dd[1] * w.y + dd[2] * w.z + dd[3] * w.w) * dactf<ACT>(h);
        }
        tc::store_bf16x8(u4 + H2_HI, u4 + H2_LO, c8, row, t8);
        tc::store_bf16x8(u4 + D2_HI, u4 + D2_LO, c8, row, d2);
      }
      tc::store_bf16x8(u4 + DO_HI, u4 + DO_LO, 0, row, dd);
      publish();
      if (warp == 2 && tc::elect_one()) {
        tc::tc_fence_after();
        gemm_rows_stacked<64>(tmem + C_W2, sa(H1_HI), sa(D2_HI), started ? 1u : 0u);   // [H1|H2 hi, lo]^T [D2 hi|lo | dout hi|lo]
        gemm_k<2, NB>(tmem + C_ACC, sa(D2_HI), sa(D2_LO), sa(WTB));
        tc::mma_commit(bar_f);
      }
      wait_f();
      // ---- delta 1 -> D1 W1^T, dW1^T --------------------------------------------------------------------------------
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
      if (warp == 3 && tc::elect_one()) {
        tc::tc_fence_after();
        gemm_k<2, 16>(tmem + C_ACC, sa(D1_HI), sa(D1_LO), sa(W1T));
        tc::mma_commit(bar_f);
        gemm_rows_stacked<32>(tmem + C_W1, sa(D1_HI), sa(XA_HI), started ? 1u : 0u);
        tc::mma_commit(bar_w);
      }
      started = 1;
      pending_w = 1;
      wait_f();
      float* const dxs = dxx + (i & 1) * (3 * TR);
      {
        float dx[8];
        load_acc<16, 8>(lane_base + C_ACC, dx);
        if (role == 0) {
          dxs[row] = dx[2];                              // d / d hS through the projected player's network
        } else {
          dxs[TR + row] = dx[2];                         // d / d S
          dxs[2 * TR + row] = dx[4];                     // d / d hS through the individual player's network
        }
      }
      tc::tc_fence_before();
      __syncthreads();
      hSbar += dxs[row] + dxs[2 * TR + row];
      Sbar += dxs[TR + row];
    }
    if (a.scheme == SCH_GLOBAL && role == 0) { y0h += hYbar * msk; y0i += Ybar * msk; }
  }
  // ---- flush: TMEM weight gradients -> scratch -> gradient vector (external flat layout) -> this CTA's row of gpart --------
  if (pending_w) { tc::mbar_wait(bar_w, phase_w); phase_w ^= 1; pending_w = 0; }
  tc::tc_fence_after();
  const float t0 = block_sum2(y0h, red);
  const float t1 = block_sum2(y0i, red);
  __syncthreads();
  float* const S = rf;                                 // this role's (dead) operand tiles: [128 lanes][SW]
  float* const sg = smem + 9216;                       // inside role 0's region, behind its scratch
  for (int e = threadIdx.x; e < a.P; e += kT) sg[e] = 0.0f;
  float* const g = sg + net.ext_off;
  const int o2 = nin * H + H, o3 = o2 + H * H + H;
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
    if (started) {
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        if (pass == 0 && c8 >= 4) break;
        float v[8];
        tc::tmem_ld8(lane_base + (pass == 0 ? C_W1 : C_W2) + 8 * c8, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) S[row * SW + 8 * c8 + q] = v[q];
      }
    }
    __syncthreads();
    if (started && pass == 0) {                        // lanes: D1 hi j, D1 lo 24 + j; columns: X hi i, X lo 16 + i
      for (int e = row; e < (nin + 1) * H; e += TR) {
        const int i = e / H, j = e % H;                // i = nin: b1
        g[e] = S[j * SW + i] + S[j * SW + 16 + i] + S[(24 + j) * SW + i];
      }
    } else if (started) {                              // lanes: H1 hi 0..23, H2 hi 24..47, H1 lo 48..71, H2 lo 72..95
      for (int e = row; e < (H + 1) * H; e += TR) {    // columns: D2 hi 0..23, D2 lo 24..47, dout hi 48..55, dout lo 56..63
        const int k = e / H, j = e % H;                // k = H: b2
        g[o2 + e] = S[k * SW + j] + S[k * SW + 24 + j] + S[(48 + k) * SW + j];
      }
      for (int e = row; e < (H + 1) * nout; e += TR) {
        const int k = e / nout, o = e % nout;          // k = H: b3
        g[o3 + e] = S[(24 + k) * SW + 48 + o] + S[(24 + k) * SW + 56 + o] + S[(72 + k) * SW + 48 + o];
      }
    }
  }
  if (a.scheme == SCH_GLOBAL && threadIdx.x == 0) { sg[a.y0_off] = t0; sg[a.y0_off + 1] = t1; }
  tc::tc_fence_before();
  __syncthreads();
  float* const grow = a.gpart + (size_t)blockIdx.x * a.P;
  for (int e = threadIdx.x; e < a.P; e += kT) grow[e] = sg[e];
  if (threadIdx.x < 32) tc::tmem_dealloc(*tslot, NCOLS);
}

}  // namespace mtc

size_t mfg_tc_smem(bool backward) { return sizeof(float) * (size_t)(backward ? mtc::b::SMEM_FLOATS : mtc::f::SMEM_FLOATS); }

int launch_mfg_tc(const MFGArgs& a, int grid, bool backward, cudaStream_t st) {
  if (a.netA.H > 22 || a.netB.H > 22 || a.netA.nout > 4 || a.netB.nout > 4 || a.netA.act != a.netB.act) {
    set_error("tcgen05 MFG kernels: need H <= 22, nout <= 4 and one activation for both networks");
    return -1;
  }
  const size_t smem = mfg_tc_smem(backward);
  const bool tanh_ = a.netA.act == ACT_TANH;
  if (!backward) {
    auto kern = tanh_ ? mtc::mfg_forward_tc<ACT_TANH> : mtc::mfg_forward_tc<ACT_RELU>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, mtc::kT, smem, st>>>(a);
  } else {
    auto kern = tanh_ ? mtc::mfg_backward_tc<ACT_TANH> : mtc::mfg_backward_tc<ACT_RELU>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, mtc::kT, smem, st>>>(a);
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fbsdej

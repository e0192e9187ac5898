// tcgen05 versions of the one-network MLP evaluation used by the compensator-free (`Reg`) solvers at large batch:
// the H x H layer (forward), its transpose (input gradient) and all three weight-gradient GEMMs run on the
// 5th-generation tensor cores with accumulators in TMEM; the thin first / last layers stay on FFMA.
//
//   TcForward  : y = W3 . act(W2 . act(W1 x))          layer 2 = 3xTF32 (fp32-grade: loss / trajectories keep 1e-5 parity)
//   TcBackward : recompute, delta pass, dL/dx, and dW1 / dW2 / dW3 accumulated in TMEM over ALL time steps and tiles
//                of the CTA (read once at kernel end); every GEMM bf16x3 (hi/lo split, ~5e-6 relative)
//
// Row r of the tile = thread r = TMEM lane r (M = 128, cta_group::1).  Operand tiles live in shared memory as 16-byte
// chunks [feature / c][row][c] (c = 4 fp32 or 8 bf16): K-major canonical when features are K, MN-major canonical when
// rows are K (tc.cuh).  The constant-1 feature carries the biases through the GEMMs.
#pragma once
#include "tc.cuh"
#include "tile_mlp.cuh"

namespace fbsdej {

// ---- forward ---------------------------------------------------------------------------------------------------
template <int NIN1>   // NIN1 = nin + 1 (inputs incl. the constant-1 feature)
struct TcForward {
  static constexpr int HP = 24;
  static constexpr int K1 = (NIN1 + 3) & ~3;
  // shared-memory carve-up (floats)
  static constexpr int OFF_AHI = 0, OFF_ALO = 3072, OFF_BHI = 6144, OFF_BLO = 6144 + 768, OFF_W1 = 6144 + 1536,
                       OFF_W3 = OFF_W1 + K1 * HP, OFF_BAR = OFF_W3 + HP, FLOATS = OFF_BAR + 8;
  float* sm;
  uint64_t* bar;
  uint32_t tmem, phase;
  int H, act;

  // all threads; theta = external flat vector, rt = the network (nin = NIN1 - 1, nout = 1)
  __device__ void init(float* smem, const float* __restrict__ theta, const NetRt& rt) {
    sm = smem; H = rt.H; act = rt.act; phase = 0;
    bar = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 2);
    for (int i = threadIdx.x; i < FLOATS; i += blockDim.x) sm[i] = 0.0f;
    __syncthreads();
    const int nin = rt.nin;
    const float* __restrict__ th = theta + rt.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H;
    for (int e = threadIdx.x; e <= n5; e += blockDim.x) {
      const float v = th[e];
      if (e < n2) {                                   // W1[i][j] and b1[j] -> W1 rows (row nin = bias)
        const int i = e < n1 ? e / H : nin, j = e < n1 ? e % H : e - n1;
        sm[OFF_W1 + i * HP + j] = v;
      } else if (e < n4) {                            // W2[k][j], b2[j] (k = H) -> B operand [k/4][n = j][k%4], hi / lo
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        float hi, lo;
        tc::split_tf32(v, hi, lo);
        sm[OFF_BHI + ((k >> 2) * 32 + j) * 4 + (k & 3)] = hi;
        sm[OFF_BLO + ((k >> 2) * 32 + j) * 4 + (k & 3)] = lo;
      } else {                                        // W3[k][0], b3 -> W3 vector (index H = bias)
        sm[OFF_W3 + (e < n5 ? e - n4 : H)] = v;
      }
    }
    if (threadIdx.x < 32) tc::tmem_alloc(tslot, 32);
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tmem = *tslot;
  }
  __device__ void finish() {
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 32);
  }

  // in[0 .. NIN1): inputs of this thread's row incl. the trailing 1.  Returns the single network output.
  __device__ __forceinline__ float eval(const float (&in)[K1]) {
    const int row = threadIdx.x;
    float a[HP];
#pragma unroll
    for (int j = 0; j < HP; ++j) a[j] = 0.0f;
#pragma unroll
    for (int k = 0; k < NIN1; ++k) axpy_row<HP>(a, in[k], sm + OFF_W1 + k * HP);
#pragma unroll
    for (int c = 0; c < HP / 4; ++c) {
      float hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 4 * c + i;
        const float h = (j < H) ? act_fn(a[j], act) : ((j == H) ? 1.0f : 0.0f);
        tc::split_tf32(h, hi[i], lo[i]);
      }
      st4(sm + OFF_AHI + (c * TR + row) * 4, make_float4(hi[0], hi[1], hi[2], hi[3]));
      st4(sm + OFF_ALO + (c * TR + row) * 4, make_float4(lo[0], lo[1], lo[2], lo[3]));
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc::tc_fence_after();
      const uint32_t id = tc::idesc_tf32(128, 32, false, false);
      const uint32_t ah = tc::smem_u32(sm + OFF_AHI), al = tc::smem_u32(sm + OFF_ALO);
      const uint32_t bh = tc::smem_u32(sm + OFF_BHI), bl = tc::smem_u32(sm + OFF_BLO);
#pragma unroll
      for (int s = 0; s < 3; ++s) {                   // K = 24 = 3 x 8
        const uint64_t dah = tc::smem_desc(ah + s * 4096, 2048, 128), dal = tc::smem_desc(al + s * 4096, 2048, 128);
        const uint64_t dbh = tc::smem_desc(bh + s * 1024, 512, 128), dbl = tc::smem_desc(bl + s * 1024, 512, 128);
        tc::mma_tf32(tmem, dah, dbh, id, s > 0 ? 1u : 0u);
        tc::mma_tf32(tmem, dal, dbh, id, 1u);
        tc::mma_tf32(tmem, dah, dbl, id, 1u);
      }
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::tc_fence_after();
    const uint32_t lane_base = tmem + ((uint32_t)(row & ~31) << 16);
    float y = sm[OFF_W3 + H];
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float v[8];
      tc::tmem_ld8(lane_base + 8 * c8, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = 8 * c8 + i;
        if (j < H) y = fmaf(act_fn(v[i], act), sm[OFF_W3 + j], y);
      }
    }
    return y;
  }
};

// ---- backward --------------------------------------------------------------------------------------------------
template <int NIN1>
struct TcBackward {
  static constexpr int HP = 24;
  static constexpr int K1 = (NIN1 + 3) & ~3;
  static_assert(NIN1 <= 16, "x tile holds 16 features");
  // uint4 offsets of the bf16 tiles ([chunk][128 rows]; hi then lo)
  static constexpr int X_HI = 0, X_LO = 256, H1_HI = 512, H1_LO = 896, H2_HI = 1280, H2_LO = 1664, D1_HI = 2048, D1_LO = 2432,
                       D2_HI = 2816, D2_LO = 3200, DO_HI = 3584, DO_LO = 3712, W2_HI = 3840, W2_LO = 3968, WT_HI = 4096,
                       WT_LO = 4224, U4_END = 4352;
  // float region after the uint4 region
  static constexpr int OFF_W1 = U4_END * 4, OFF_W3 = OFF_W1 + K1 * HP, OFF_BAR = OFF_W3 + HP, FLOATS = OFF_BAR + 8;
  // TMEM columns
  static constexpr uint32_t C_ACC = 0, C_W2 = 32, C_W1 = 64, C_W3 = 96, NCOLS = 128;
  float* sm;
  uint4* u4;
  uint64_t* bar_f;
  uint64_t* bar_w;
  uint32_t tmem, phase_f, phase_w, pending_w, started;
  int H, act;

  __device__ void init(float* smem, const float* __restrict__ theta, const NetRt& rt) {
    sm = smem; u4 = reinterpret_cast<uint4*>(smem); H = rt.H; act = rt.act;
    phase_f = phase_w = pending_w = started = 0;
    bar_f = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    bar_w = bar_f + 1;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 4);
    for (int i = threadIdx.x; i < FLOATS; i += blockDim.x) sm[i] = 0.0f;
    __syncthreads();
    const int nin = rt.nin;
    const float* __restrict__ th = theta + rt.ext_off;
    const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H;
    unsigned short* w2h = reinterpret_cast<unsigned short*>(u4 + W2_HI);
    unsigned short* w2l = reinterpret_cast<unsigned short*>(u4 + W2_LO);
    unsigned short* wth = reinterpret_cast<unsigned short*>(u4 + WT_HI);
    unsigned short* wtl = reinterpret_cast<unsigned short*>(u4 + WT_LO);
    for (int e = threadIdx.x; e <= n5; e += blockDim.x) {
      const float v = th[e];
      if (e < n2) {
        const int i = e < n1 ? e / H : nin, j = e < n1 ? e % H : e - n1;
        sm[OFF_W1 + i * HP + j] = v;
      } else if (e < n4) {
        const int k = e < n3 ? (e - n2) / H : H, j = e < n3 ? (e - n2) % H : e - n3;
        uint32_t hi, lo;
        tc::split_bf16(v, hi, lo);
        // forward B operand: N = j (output unit), K = k (input unit incl. the bias row H): [k/8][n][k%8]
        w2h[((k >> 3) * 32 + j) * 8 + (k & 7)] = (unsigned short)hi;
        w2l[((k >> 3) * 32 + j) * 8 + (k & 7)] = (unsigned short)lo;
        if (k < H) {                                  // input-gradient B operand: N = k, K = j
          wth[((j >> 3) * 32 + k) * 8 + (j & 7)] = (unsigned short)hi;
          wtl[((j >> 3) * 32 + k) * 8 + (j & 7)] = (unsigned short)lo;
        }
      } else {
        sm[OFF_W3 + (e < n5 ? e - n4 : H)] = v;
      }
    }
    if (threadIdx.x < 32) tc::tmem_alloc(tslot, NCOLS);
    if (threadIdx.x == 0) { tc::mbar_init(bar_f, 1); tc::mbar_init(bar_w, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tmem = *tslot;
  }

  __device__ __forceinline__ void wait_f() {
    tc::mbar_wait(bar_f, phase_f);
    phase_f ^= 1;
    tc::tc_fence_after();
  }
  __device__ __forceinline__ void publish() {      // smem tiles written by all threads -> MMA issue by thread 0
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
  }
  // D[cols 0..31] = A(K-major bf16x3, 128 x 32) * B([N = 32][K = 32]);  thread 0 only
  __device__ __forceinline__ void gemm_k(int a_hi, int a_lo, int b_hi, int b_lo) {
    const uint32_t id = tc::idesc_bf16(128, 32, false, false);
    const uint32_t ah = tc::smem_u32(u4 + a_hi), al = tc::smem_u32(u4 + a_lo);
    const uint32_t bh = tc::smem_u32(u4 + b_hi), bl = tc::smem_u32(u4 + b_lo);
#pragma unroll
    for (int s = 0; s < 2; ++s) {                   // K = 32 = 2 x 16
      const uint64_t dah = tc::smem_desc(ah + s * 4096, 2048, 128), dal = tc::smem_desc(al + s * 4096, 2048, 128);
      const uint64_t dbh = tc::smem_desc(bh + s * 1024, 512, 128), dbl = tc::smem_desc(bl + s * 1024, 512, 128);
      tc::mma_bf16(tmem + C_ACC, dah, dbh, id, s > 0 ? 1u : 0u);
      tc::mma_bf16(tmem + C_ACC, dal, dbh, id, 1u);
      tc::mma_bf16(tmem + C_ACC, dah, dbl, id, 1u);
    }
  }
  // D[col0 ..] += sum over the 128 rows of A^T B (both MN-major bf16x3);  thread 0 only
  template <int N>
  __device__ __forceinline__ void gemm_rows(uint32_t col0, int a_hi, int a_lo, int b_hi, int b_lo, uint32_t acc0) {
    const uint32_t id = tc::idesc_bf16(128, N, true, true);
    const uint32_t ah = tc::smem_u32(u4 + a_hi), al = tc::smem_u32(u4 + a_lo);
    const uint32_t bh = tc::smem_u32(u4 + b_hi), bl = tc::smem_u32(u4 + b_lo);
#pragma unroll
    for (int s = 0; s < 8; ++s) {                   // 128 rows = 8 x 16
      const uint64_t dah = tc::smem_desc(ah + s * 256, 128, 2048), dal = tc::smem_desc(al + s * 256, 128, 2048);
      const uint64_t dbh = tc::smem_desc(bh + s * 256, 128, 2048), dbl = tc::smem_desc(bl + s * 256, 128, 2048);
      tc::mma_bf16(tmem + col0, dah, dbh, id, (s > 0) ? 1u : acc0);
      tc::mma_bf16(tmem + col0, dal, dbh, id, 1u);
      tc::mma_bf16(tmem + col0, dah, dbl, id, 1u);
    }
  }
  __device__ __forceinline__ void load_acc(float (&v)[HP]) {
    const uint32_t lane_base = tmem + ((uint32_t)(threadIdx.x & ~31) << 16) + C_ACC;
#pragma unroll
    for (int c8 = 0; c8 < 3; ++c8) {
      float t8[8];
      tc::tmem_ld8(lane_base + 8 * c8, t8);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * c8 + i] = t8[i];
    }
  }

  // One row: inputs in[0 .. NIN1) (trailing 1 included), adjoint dout of the single output (0 for rows that do not
  // count).  Returns dL/d in[i] in dx[i], i < NIN1 - 1; accumulates the weight gradients of the whole tile in TMEM.
  __device__ __forceinline__ void step(const float (&in)[K1], float dout, float (&dx)[K1]) {
    const int row = threadIdx.x;
    if (pending_w) {                                 // the previous step's weight-gradient MMAs still read the tiles
      tc::mbar_wait(bar_w, phase_w);
      phase_w ^= 1;
      pending_w = 0;
    }
    float a[HP];
    {  // layer 1 (fp32 FFMA) -> h1, stored with the input row
#pragma unroll
      for (int j = 0; j < HP; ++j) a[j] = 0.0f;
#pragma unroll
      for (int k = 0; k < NIN1; ++k) axpy_row<HP>(a, in[k], sm + OFF_W1 + k * HP);
#pragma unroll
      for (int j = 0; j < HP; ++j) a[j] = (j < H) ? act_fn(a[j], act) : ((j == H) ? 1.0f : 0.0f);
      float xin[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) xin[k] = (k < K1) ? in[k < K1 ? k : 0] : 0.0f;
      tc::store_bf16x8(u4 + X_HI, u4 + X_LO, 0, row, xin);
      tc::store_bf16x8(u4 + X_HI, u4 + X_LO, 1, row, xin + 8);
#pragma unroll
      for (int c = 0; c < 3; ++c) tc::store_bf16x8(u4 + H1_HI, u4 + H1_LO, c, row, a + 8 * c);
    }
    publish();
    if (threadIdx.x == 0) {
      tc::tc_fence_after();
      gemm_k(H1_HI, H1_LO, W2_HI, W2_LO);
      tc::mma_commit(bar_f);
    }
    wait_f();
    {  // h2, delta 2
      load_acc(a);
      float d2[HP];
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        const float h = (j < H) ? act_fn(a[j], act) : ((j == H) ? 1.0f : 0.0f);
        a[j] = h;
        d2[j] = (j < H) ? dout * sm[OFF_W3 + j] * dact_fn(h, act) : 0.0f;
      }
      float dd[8] = {dout, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      tc::store_bf16x8(u4 + DO_HI, u4 + DO_LO, 0, row, dd);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tc::store_bf16x8(u4 + H2_HI, u4 + H2_LO, c, row, a + 8 * c);
        tc::store_bf16x8(u4 + D2_HI, u4 + D2_LO, c, row, d2 + 8 * c);
      }
    }
    publish();
    if (threadIdx.x == 0) {
      tc::tc_fence_after();
      gemm_k(D2_HI, D2_LO, WT_HI, WT_LO);
      tc::mma_commit(bar_f);
    }
    wait_f();
    {  // delta 1 = (d2 W2^T) .* act'(h1);  h1 is re-read from its bf16 hi + lo chunks
      load_acc(a);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 hh = u4[H1_HI + c * 128 + row], hl = u4[H1_LO + c * 128 + row];
        const uint32_t wh[4] = {hh.x, hh.y, hh.z, hh.w}, wl[4] = {hl.x, hl.y, hl.z, hl.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * c + i;
          const uint32_t bh = (i & 1) ? (wh[i >> 1] & 0xFFFF0000u) : (wh[i >> 1] << 16);
          const uint32_t bl = (i & 1) ? (wl[i >> 1] & 0xFFFF0000u) : (wl[i >> 1] << 16);
          const float h = __uint_as_float(bh) + __uint_as_float(bl);
          a[j] = (j < H) ? a[j] * dact_fn(h, act) : 0.0f;
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) tc::store_bf16x8(u4 + D1_HI, u4 + D1_LO, c, row, a + 8 * c);
      // dL/dx_i = sum_j W1[i][j] d1[j]
#pragma unroll
      for (int i = 0; i < NIN1 - 1; ++i) {
        const float* __restrict__ w = sm + OFF_W1 + i * HP;
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; ++j4) {
          const float4 wv = ld4(w + 4 * j4);
          s0 = fmaf(wv.x, a[4 * j4], s0); s1 = fmaf(wv.y, a[4 * j4 + 1], s1);
          s0 = fmaf(wv.z, a[4 * j4 + 2], s0); s1 = fmaf(wv.w, a[4 * j4 + 3], s1);
        }
        dx[i] = s0 + s1;
      }
    }
    publish();
    if (threadIdx.x == 0) {
      tc::tc_fence_after();
      const uint32_t acc0 = started ? 1u : 0u;
      gemm_rows<32>(C_W2, H1_HI, H1_LO, D2_HI, D2_LO, acc0);     // dW2[k][j] = sum_r h1[r][k] d2[r][j]   (k = H: b2)
      gemm_rows<32>(C_W1, X_HI, X_LO, D1_HI, D1_LO, acc0);       // dW1[i][j] = sum_r x[r][i] d1[r][j]    (i = nin: b1)
      gemm_rows<16>(C_W3, H2_HI, H2_LO, DO_HI, DO_LO, acc0);     // dW3[k]    = sum_r h2[r][k] dout[r]    (k = H: b3)
      tc::mma_commit(bar_w);
    }
    started = 1;
    pending_w = 1;
  }

  // After the last step: weight gradients TMEM -> sg[ext_off ...] (external flat layout).  All threads call.
  __device__ void flush(float* __restrict__ sg, const NetRt& rt) {
    if (pending_w) { tc::mbar_wait(bar_w, phase_w); phase_w ^= 1; pending_w = 0; }
    tc::tc_fence_after();
    if (threadIdx.x < 32 && started) {
      const int k = threadIdx.x, nin = rt.nin;
      float v[8];
      float* g = sg + rt.ext_off;
      const int o2 = nin * H + H, o3 = o2 + H * H + H;
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        tc::tmem_ld8(tmem + C_W1 + 8 * c8, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * c8 + i;
          if (k <= nin && j < H) g[k < nin ? k * H + j : nin * H + j] = v[i];
        }
        tc::tmem_ld8(tmem + C_W2 + 8 * c8, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * c8 + i;
          if (k <= H && j < H) g[k < H ? o2 + k * H + j : o2 + H * H + j] = v[i];
        }
      }
      tc::tmem_ld8(tmem + C_W3, v);
      tc::tmem_ld_wait();
      if (k <= H) g[k < H ? o3 + k : o3 + H] = v[0];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, NCOLS);
  }
};

}  // namespace fbsdej

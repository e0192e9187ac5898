// Self-test of the tcgen05 plumbing (tc.cuh): two tiny GEMMs on one CTA, results returned to the caller.
//   test 0: D[128][32] = A[128][24] * B[24][32]          K-major operands  (forward / input-gradient GEMM shape)
//   test 1: D[m][n]    = sum_r P[r][m] * Q[r][n], r<128  MN-major operands (weight-gradient GEMM shape), m,n < 24
// Both with the 3xTF32 split.  Used by tests/test_tc_gpu.py before the fused kernels rely on the same descriptors.
#include "common.cuh"
#include "tc.cuh"

namespace fbsdej {

constexpr int ST_TR = 128;

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          const float* __restrict__ P, const float* __restrict__ Q,
                                                          float* __restrict__ out0, float* __restrict__ out1) {
  extern __shared__ __align__(1024) float sm[];
  // tiles: [6 chunks][128 rows][4]; placed first so that the M = 128 MN-major reads (32 chunks x 2 KB) stay inside smem
  float* a_hi = sm;                 // 3072 floats
  float* a_lo = a_hi + 3072;
  float* q_hi = a_lo + 3072;
  float* q_lo = q_hi + 3072;
  float* b_hi = q_lo + 3072;        // [6 chunks][32 n][4]
  float* b_lo = b_hi + 768;
  // (>= 64 KB of slack follows b_lo: dynamic smem size chosen by the launcher)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int r = threadIdx.x, warp = r >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
  if (r == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  uint32_t phase = 0;
  for (int test = 0; test < 2; ++test) {
    const float* src = test == 0 ? A : P;
    for (int c = 0; c < 6; ++c) {
      float hi[4], lo[4];
      for (int i = 0; i < 4; ++i) tc::split_tf32(src[r * 24 + 4 * c + i], hi[i], lo[i]);
      st4(a_hi + (c * ST_TR + r) * 4, make_float4(hi[0], hi[1], hi[2], hi[3]));
      st4(a_lo + (c * ST_TR + r) * 4, make_float4(lo[0], lo[1], lo[2], lo[3]));
      if (test == 1) {
        for (int i = 0; i < 4; ++i) tc::split_tf32(Q[r * 24 + 4 * c + i], hi[i], lo[i]);
        st4(q_hi + (c * ST_TR + r) * 4, make_float4(hi[0], hi[1], hi[2], hi[3]));
        st4(q_lo + (c * ST_TR + r) * 4, make_float4(lo[0], lo[1], lo[2], lo[3]));
      }
    }
    if (test == 0) {
      for (int e = r; e < 24 * 32; e += 128) {
        const int k = e / 32, n = e % 32;
        float hi, lo;
        tc::split_tf32(B[k * 32 + n], hi, lo);
        b_hi[((k >> 2) * 32 + n) * 4 + (k & 3)] = hi;
        b_lo[((k >> 2) * 32 + n) * 4 + (k & 3)] = lo;
      }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (r == 0) {
      tc::tc_fence_after();
      if (test == 0) {
        const uint32_t id = tc::idesc_tf32(128, 32, false, false);
        uint32_t acc = 0;
        for (int s = 0; s < 3; ++s) {       // K = 24 = 3 x 8
          const uint64_t ah = tc::smem_desc(tc::smem_u32(a_hi) + s * 4096, 2048, 128);
          const uint64_t al = tc::smem_desc(tc::smem_u32(a_lo) + s * 4096, 2048, 128);
          const uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi) + s * 1024, 512, 128);
          const uint64_t bl = tc::smem_desc(tc::smem_u32(b_lo) + s * 1024, 512, 128);
          tc::mma_tf32(tm, ah, bh, id, acc); acc = 1;
          tc::mma_tf32(tm, al, bh, id, 1);
          tc::mma_tf32(tm, ah, bl, id, 1);
        }
      } else {
        const uint32_t id = tc::idesc_tf32(128, 32, true, true);
        uint32_t acc = 0;
        for (int s = 0; s < 16; ++s) {      // K = 128 rows = 16 x 8
          const uint64_t ph = tc::smem_desc(tc::smem_u32(a_hi) + s * 128, 128, 2048);
          const uint64_t pl = tc::smem_desc(tc::smem_u32(a_lo) + s * 128, 128, 2048);
          const uint64_t qh = tc::smem_desc(tc::smem_u32(q_hi) + s * 128, 128, 2048);
          const uint64_t ql = tc::smem_desc(tc::smem_u32(q_lo) + s * 128, 128, 2048);
          tc::mma_tf32(tm + 32, ph, qh, id, acc); acc = 1;
          tc::mma_tf32(tm + 32, pl, qh, id, 1);
          tc::mma_tf32(tm + 32, ph, ql, id, 1);
        }
      }
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    tc::tc_fence_after();
    float* out = test == 0 ? out0 : out1;
    for (int c8 = 0; c8 < 4; ++c8) {
      float v[8];
      tc::tmem_ld8(tm + ((uint32_t)(32 * warp) << 16) + (test == 0 ? 0 : 32) + 8 * c8, v);
      tc::tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[r * 32 + 8 * c8 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tc::tmem_dealloc(tm, 64);
}

// bf16x3 variants: test 2: out2[128][32] = A[128][24(->32)] * B[24][32] (K-major); test 3: out3[m][n] = sum_r P[r][m] Q[r][n]
// (MN-major).  Tiles [feature/8][row][8 bf16]: one byte layout serves both major-nesses.
__global__ void __launch_bounds__(128) tc_selftest_bf16_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               const float* __restrict__ P, const float* __restrict__ Q,
                                                               float* __restrict__ out2, float* __restrict__ out3) {
  extern __shared__ __align__(1024) float sm[];
  uint4* a_hi = reinterpret_cast<uint4*>(sm);        // 4 chunks x 128 rows (chunk 3 zero)
  uint4* a_lo = a_hi + 512;
  uint4* q_hi = a_lo + 512;
  uint4* q_lo = q_hi + 512;
  uint4* b_hi = q_lo + 512;                          // [K/8 = 4][N = 32] uint4 (8 bf16 along K)
  uint4* b_lo = b_hi + 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int r = threadIdx.x, warp = r >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
  if (r == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  uint32_t phase = 0;
  for (int test = 0; test < 2; ++test) {
    const float* src = test == 0 ? A : P;
    for (int c = 0; c < 4; ++c) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = (8 * c + i < 24) ? src[r * 24 + 8 * c + i] : 0.0f;
      tc::store_bf16x8(a_hi, a_lo, c, r, v);
      if (test == 1) {
        for (int i = 0; i < 8; ++i) v[i] = (8 * c + i < 24) ? Q[r * 24 + 8 * c + i] : 0.0f;
        tc::store_bf16x8(q_hi, q_lo, c, r, v);
      }
    }
    if (test == 0) {
      unsigned short* bh = reinterpret_cast<unsigned short*>(b_hi);
      unsigned short* bl = reinterpret_cast<unsigned short*>(b_lo);
      for (int e = r; e < 32 * 32; e += 128) {
        const int k = e / 32, n = e % 32;
        uint32_t hi, lo;
        tc::split_bf16(k < 24 ? B[k * 32 + n] : 0.0f, hi, lo);
        bh[((k >> 3) * 32 + n) * 8 + (k & 7)] = (unsigned short)hi;
        bl[((k >> 3) * 32 + n) * 8 + (k & 7)] = (unsigned short)lo;
      }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (r == 0) {
      tc::tc_fence_after();
      if (test == 0) {
        const uint32_t id = tc::idesc_bf16(128, 32, false, false);
        uint32_t acc = 0;
        for (int s = 0; s < 2; ++s) {       // K = 32 = 2 x 16 (two 16-byte chunks per step)
          const uint64_t ah = tc::smem_desc(tc::smem_u32(a_hi) + s * 4096, 2048, 128);
          const uint64_t al = tc::smem_desc(tc::smem_u32(a_lo) + s * 4096, 2048, 128);
          const uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi) + s * 1024, 512, 128);
          const uint64_t bl = tc::smem_desc(tc::smem_u32(b_lo) + s * 1024, 512, 128);
          tc::mma_bf16(tm, ah, bh, id, acc); acc = 1;
          tc::mma_bf16(tm, al, bh, id, 1);
          tc::mma_bf16(tm, ah, bl, id, 1);
        }
      } else {
        const uint32_t id = tc::idesc_bf16(128, 32, true, true);
        uint32_t acc = 0;
        for (int s = 0; s < 8; ++s) {       // K = 128 rows = 8 x 16 (two 8-row groups per step, 128 B apart)
          const uint64_t ph = tc::smem_desc(tc::smem_u32(a_hi) + s * 256, 128, 2048);
          const uint64_t pl = tc::smem_desc(tc::smem_u32(a_lo) + s * 256, 128, 2048);
          const uint64_t qh = tc::smem_desc(tc::smem_u32(q_hi) + s * 256, 128, 2048);
          const uint64_t ql = tc::smem_desc(tc::smem_u32(q_lo) + s * 256, 128, 2048);
          tc::mma_bf16(tm + 32, ph, qh, id, acc); acc = 1;
          tc::mma_bf16(tm + 32, pl, qh, id, 1);
          tc::mma_bf16(tm + 32, ph, ql, id, 1);
        }
      }
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    tc::tc_fence_after();
    float* out = test == 0 ? out2 : out3;
    for (int c8 = 0; c8 < 4; ++c8) {
      float v[8];
      tc::tmem_ld8(tm + ((uint32_t)(32 * warp) << 16) + (test == 0 ? 0 : 32) + 8 * c8, v);
      tc::tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[r * 32 + 8 * c8 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tc::tmem_dealloc(tm, 64);
}

// A operand in TMEM (tcgen05.st by the row's own thread), B in shared memory; plus an M = 64 MN-major GEMM whose
// accumulator is dumped from all 128 lanes so that the test can recover the row -> lane map.
//   out_a[0] = A * B, 3xTF32, A from TMEM      out_a[1] = A * B, bf16x3, A from TMEM      out_m64 = all lanes x 32 columns
__global__ void __launch_bounds__(128) tc_selftest_tmemA_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                const float* __restrict__ P, const float* __restrict__ Q,
                                                                float* __restrict__ out_a, float* __restrict__ out_m64) {
  extern __shared__ __align__(1024) float sm[];
  float* bt_hi = sm;                                   // tf32 B operand [6 chunks][32 n][4]
  float* bt_lo = bt_hi + 768;
  uint4* bb_hi = reinterpret_cast<uint4*>(bt_lo + 768);  // bf16 B operand [4 chunks][32 n] uint4
  uint4* bb_lo = bb_hi + 128;
  uint4* p_hi = bb_lo + 128;                           // bf16 tiles [4 chunks][128 rows] uint4
  uint4* p_lo = p_hi + 512;
  uint4* q_hi = p_lo + 512;
  uint4* q_lo = q_hi + 512;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int r = threadIdx.x, warp = r >> 5;
  for (int i = r; i < 2 * 768 + 4 * 256 + 4 * 4 * 512 + 8192; i += 128) sm[i] = 0.0f;
  __syncthreads();
  if (warp == 0) tc::tmem_alloc(&tmem_base, 128);
  if (r == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  for (int e = r; e < 24 * 32; e += 128) {
    const int k = e / 32, n = e % 32;
    float hi, lo;
    tc::split_tf32(B[k * 32 + n], hi, lo);
    bt_hi[((k >> 2) * 32 + n) * 4 + (k & 3)] = hi;
    bt_lo[((k >> 2) * 32 + n) * 4 + (k & 3)] = lo;
    uint32_t h16, l16;
    tc::split_bf16(B[k * 32 + n], h16, l16);
    reinterpret_cast<unsigned short*>(bb_hi)[((k >> 3) * 32 + n) * 8 + (k & 7)] = (unsigned short)h16;
    reinterpret_cast<unsigned short*>(bb_lo)[((k >> 3) * 32 + n) * 8 + (k & 7)] = (unsigned short)l16;
  }
  for (int c = 0; c < 4; ++c) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = (8 * c + i < 24) ? P[r * 24 + 8 * c + i] : 0.0f;
    tc::store_bf16x8(p_hi, p_lo, c, r, v);
    for (int i = 0; i < 8; ++i) v[i] = (8 * c + i < 24) ? Q[r * 24 + 8 * c + i] : 0.0f;
    tc::store_bf16x8(q_hi, q_lo, c, r, v);
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base, lane_base = tm + ((uint32_t)(32 * warp) << 16);
  uint32_t phase = 0;
  // ---- test 0: tf32, A hi at columns 32..55, A lo at 64..87 -----------------------------------------------------
  for (int c8 = 0; c8 < 3; ++c8) {
    uint32_t h[8], l[8];
    for (int i = 0; i < 8; ++i) {
      float hi, lo;
      tc::split_tf32(A[r * 24 + 8 * c8 + i], hi, lo);
      h[i] = __float_as_uint(hi); l[i] = __float_as_uint(lo);
    }
    tc::tmem_st8(lane_base + 32 + 8 * c8, h);
    tc::tmem_st8(lane_base + 64 + 8 * c8, l);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  if (r == 0) {
    tc::tc_fence_after();
    const uint32_t id = tc::idesc_tf32(128, 32, false, false);
    for (int s = 0; s < 3; ++s) {
      const uint64_t bh = tc::smem_desc(tc::smem_u32(bt_hi) + s * 1024, 512, 128), bl = tc::smem_desc(tc::smem_u32(bt_lo) + s * 1024, 512, 128);
      tc::mma_tf32_ts(tm, tm + 32 + 8 * s, bh, id, s > 0 ? 1u : 0u);
      tc::mma_tf32_ts(tm, tm + 64 + 8 * s, bh, id, 1u);
      tc::mma_tf32_ts(tm, tm + 32 + 8 * s, bl, id, 1u);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, phase); phase ^= 1;
  tc::tc_fence_after();
  for (int c8 = 0; c8 < 4; ++c8) {
    float v[8];
    tc::tmem_ld8(lane_base + 8 * c8, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out_a[r * 32 + 8 * c8 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  // ---- test 1: bf16, K = 32 (24 real): A hi packed pairs at columns 32..47, A lo at 48..63 ----------------------------
  for (int c8 = 0; c8 < 2; ++c8) {
    uint32_t h[8], l[8];
    for (int i = 0; i < 8; ++i) {
      const int k0 = 16 * c8 + 2 * i;
      uint32_t h0, l0, h1, l1;
      tc::split_bf16(k0 < 24 ? A[r * 24 + k0] : 0.0f, h0, l0);
      tc::split_bf16(k0 + 1 < 24 ? A[r * 24 + k0 + 1] : 0.0f, h1, l1);
      h[i] = h0 | (h1 << 16); l[i] = l0 | (l1 << 16);
    }
    tc::tmem_st8(lane_base + 32 + 8 * c8, h);
    tc::tmem_st8(lane_base + 48 + 8 * c8, l);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  if (r == 0) {
    tc::tc_fence_after();
    const uint32_t id = tc::idesc_bf16(128, 32, false, false);
    for (int s = 0; s < 2; ++s) {
      const uint64_t bh = tc::smem_desc(tc::smem_u32(bb_hi) + s * 1024, 512, 128), bl = tc::smem_desc(tc::smem_u32(bb_lo) + s * 1024, 512, 128);
      tc::mma_bf16_ts(tm, tm + 32 + 8 * s, bh, id, s > 0 ? 1u : 0u);
      tc::mma_bf16_ts(tm, tm + 48 + 8 * s, bh, id, 1u);
      tc::mma_bf16_ts(tm, tm + 32 + 8 * s, bl, id, 1u);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, phase); phase ^= 1;
  tc::tc_fence_after();
  for (int c8 = 0; c8 < 4; ++c8) {
    float v[8];
    tc::tmem_ld8(lane_base + 8 * c8, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out_a[128 * 32 + r * 32 + 8 * c8 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  // ---- test 2: M = 64, MN-major bf16x3: D[m][n] = sum_r P[r][m] Q[r][n]; first zero all lanes of the accumulator --------
  {
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c8 = 0; c8 < 4; ++c8) tc::tmem_st8(lane_base + 96 + 8 * c8, z);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (r == 0) {
    tc::tc_fence_after();
    const uint32_t id = tc::idesc_bf16(64, 32, true, true);
    for (int s = 0; s < 8; ++s) {
      const uint64_t ph = tc::smem_desc(tc::smem_u32(p_hi) + s * 256, 128, 2048), pl = tc::smem_desc(tc::smem_u32(p_lo) + s * 256, 128, 2048);
      const uint64_t qh = tc::smem_desc(tc::smem_u32(q_hi) + s * 256, 128, 2048), ql = tc::smem_desc(tc::smem_u32(q_lo) + s * 256, 128, 2048);
      tc::mma_bf16(tm + 96, ph, qh, id, 1u);
      tc::mma_bf16(tm + 96, pl, qh, id, 1u);
      tc::mma_bf16(tm + 96, ph, ql, id, 1u);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, phase); phase ^= 1;
  tc::tc_fence_after();
  for (int c8 = 0; c8 < 4; ++c8) {
    float v[8];
    tc::tmem_ld8(lane_base + 96 + 8 * c8, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out_m64[r * 32 + 8 * c8 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 128);
}

int launch_tc_selftest(const float* A, const float* B, const float* P, const float* Q, float* out0, float* out1, cudaStream_t st) {
  const size_t smem = 160 * 1024;
  FB_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, st>>>(A, B, P, Q, out0, out1);
  FB_CUDA(cudaGetLastError());
  FB_CUDA(cudaFuncSetAttribute(tc_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_bf16_kernel<<<1, 128, smem, st>>>(A, B, P, Q, out0 + 128 * 32, out1 + 128 * 32);
  FB_CUDA(cudaGetLastError());
  FB_CUDA(cudaFuncSetAttribute(tc_selftest_tmemA_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_tmemA_kernel<<<1, 128, smem, st>>>(A, B, P, Q, out0 + 2 * 128 * 32, out1 + 2 * 128 * 32);
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fbsdej

// This file is linked into libfbsdej_selftest.so (test infrastructure), NOT into the product library.
namespace fbsdej {
static thread_local std::string g_selftest_err;
void set_error(const std::string& msg) { g_selftest_err = msg; }
}  // namespace fbsdej
extern "C" __attribute__((visibility("default"))) int fbsdej_selftest_tc(void* stream, const float* A, const float* B, const float* P,
                                                                        const float* Q, float* out0, float* out1) {
  if (!A || !B || !P || !Q || !out0 || !out1) return -1;
  return fbsdej::launch_tc_selftest(A, B, P, Q, out0, out1, (cudaStream_t)stream);
}
extern "C" __attribute__((visibility("default"))) const char* fbsdej_selftest_last_error(void) { return fbsdej::g_selftest_err.c_str(); }

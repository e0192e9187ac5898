// Building blocks of the tensor-core network kernels (reg_tc_kernels.cu: compensator-free pricing solvers; mfg_tc_kernels.cu:
// the two MFG networks): split-precision GEMM issue loops over the operand layouts of tc.cuh, accumulator read-back,
// activations.  One tile = 128 rows = 128 TMEM lanes; thread r owns row r.
#pragma once
#include "tc.cuh"
#include "tile_mlp.cuh"

namespace fbsdej {
namespace rtc {

constexpr int NB = 24;                        // n-rows stored per chunk of a B operand (hidden width padded to 24)

__device__ __forceinline__ void publish() {  // generic-proxy tile writes -> async proxy, then the CTA barrier
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Layer GEMM, bf16x3 in TWO MMAs per 16-wide K slice: D[:, 0 .. 2 NH) = A_hi [B_hi | B_lo] (N = 2 NH), then
// D[:, 0 .. N2) += A_lo [B_hi | ..] (N2 = NH rounded up to 16; the columns beyond NH pick up lo.lo terms, which belong to
// the exact product).  The caller adds the column blocks [0, NH) and [NH, 2 NH) when it reads the accumulator.
// A: K-major tile (128 rows); B: [k / 8][2 NH n-rows][8 bf16].
// FULL: the second MMA spans both column blocks as well, D[:, 0 .. 2 NH) += A_lo [B_hi | B_lo]: all four hi / lo cross products
// (the jump schemes, whose gradients are differences of nearly equal sums, want the lo.lo term: 2^-17 instead of 2^-16).
// (addresses in units of 16 bytes: tc::smem_desc16).  CALL UNDER `warp == k && tc::elect_one()`, not under a `lane == 0` test: in
// a lane-divergent branch ptxas wraps every tcgen05 instruction in an elect / retry loop of five instructions.
// All MMAs of the GEMM sit in ONE asm statement: the compiler wraps every asm
// statement that uses uniform registers inside a divergent branch (the single issuing lane) in an elect / retry loop of five
// instructions, and these MMAs are on the critical path between a barrier and the next wait.
template <int KS, int NH, bool FULL = false>
__device__ __forceinline__ void gemm_k16(uint32_t tmem_d, uint32_t a_hi16, uint32_t a_lo16, uint32_t b16) {
  static_assert(KS == 1 || KS == 2, "one or two 16-wide K slices");
  constexpr uint32_t id1 = tc::idesc_bf16(128, 2 * NH, false, false),
                     id2 = tc::idesc_bf16(128, FULL ? 2 * NH : (NH + 15) / 16 * 16, false, false);
  constexpr uint32_t chunk = 2 * NH * 16;
  const uint64_t dah = tc::smem_desc16(a_hi16, 2048, 128), dal = tc::smem_desc16(a_lo16, 2048, 128), db = tc::smem_desc16(b16, chunk, 128);
  if constexpr (KS == 1) {
    asm volatile(
        "{\n\t.reg .pred pf, pt;\n\t"
        "setp.ne.u32 pf, 0, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %4, pf;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, pt;\n\t}"
        :
        : "r"(tmem_d), "l"(dah), "l"(dal), "l"(db), "r"(id1), "r"(id2)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pf, pt;\n\t.reg .b64 ah1, al1, b1;\n\t"
        "setp.ne.u32 pf, 0, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
        "add.s64 ah1, %1, 256;\n\tadd.s64 al1, %2, 256;\n\tadd.s64 b1, %3, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %4, pf;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ah1, b1, %4, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], al1, b1, %5, pt;\n\t}"
        :
        : "r"(tmem_d), "l"(dah), "l"(dal), "l"(db), "r"(id1), "r"(id2), "n"(2 * chunk / 16)
        : "memory");
  }
}
// byte-address form (16-byte aligned operands)
template <int KS, int NH, bool FULL = false>
__device__ __forceinline__ void gemm_k(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b) {
  gemm_k16<KS, NH, FULL>(tmem_d, tc::addr16(a_hi), tc::addr16(a_lo), tc::addr16(b));
}
// this thread's row of a layer GEMM result: v[j] = D[j] + D[NH + j], j < NJ (NJ a multiple of 8)
template <int NH, int NJ>
__device__ __forceinline__ void load_acc(uint32_t lane_base, float (&v)[NJ]) {
#pragma unroll
  for (int c8 = 0; c8 < NJ / 8; ++c8) {
    float p8[8], q8[8];
    tc::tmem_ld8(lane_base + 8 * c8, p8);
    tc::tmem_ld8(lane_base + NH + 8 * c8, q8);
    tc::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) v[8 * c8 + q] = p8[q] + q8[q];
  }
}
// Weight-gradient GEMM over the 128 rows of the tile (rows are K, both operands MN-major).  The hi and lo copies of A
// are contiguous along M and those of B along N, so ONE MMA per 16-row slice forms all four hi/lo cross products in
// separate accumulator blocks:  D[m][n], m in [A_hi features | A_lo features], n in [B_hi | B_lo] (24 columns each).
// The three blocks that matter (hi.hi, hi.lo, lo.hi) are added when the accumulators are read at the end of the kernel.
// MM = 64 when the stacked A features fit 64 rows: the MMA then reads half the A bytes from shared memory (these small MMAs are
// bound by their operand reads: scripts/mma_microbench.cu), and accumulator row m lands in lane (m / 16) * 32 + m % 16.
__host__ __device__ constexpr int lane_of_row_m64(int m) { return (m >> 4) * 32 + (m & 15); }
template <int NN, int MM = 128>
__device__ __forceinline__ void gemm_rows_stacked16(uint32_t tmem_d, uint32_t a16, uint32_t b16, uint32_t acc0) {
  constexpr uint32_t id = tc::idesc_bf16(MM, NN, true, true);
  const uint64_t da = tc::smem_desc16(a16, 128, 2048), db = tc::smem_desc16(b16, 128, 2048);
  asm volatile(                                     // 128 rows = 8 x 16: operand s is 256 bytes (16 units) after operand s - 1
      "{\n\t.reg .pred p, pt;\n\t.reg .b64 a, b;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
      "mov.b64 a, %1;\n\tmov.b64 b, %2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, p;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "add.s64 a, a, 16;\n\tadd.s64 b, b, 16;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t}"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"(acc0)
      : "memory");
}

template <int NN, int MM = 128>
__device__ __forceinline__ void gemm_rows_stacked(uint32_t tmem_d, uint32_t a, uint32_t b, uint32_t acc0) {
  gemm_rows_stacked16<NN, MM>(tmem_d, tc::addr16(a), tc::addr16(b), acc0);
}

__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int ACT>
__device__ __forceinline__ float actf(float x) {
#if defined(FBSDEJ_ABLATE) && FBSDEJ_ABLATE == 13
  return 0.3f * x;
#endif
  return ACT == ACT_TANH ? tanh_fast(x) : fmaxf(x, 0.0f);
}
template <int ACT>
__device__ __forceinline__ float dactf(float h) { return ACT == ACT_TANH ? fmaf(-h, h, 1.0f) : (h > 0.0f ? 1.0f : 0.0f); }


// ---- forward sweeps: 3xTF32, A operand in tensor memory --------------------------------------------------------------
namespace fwd {
constexpr uint32_t C_AHI = 0, C_ALO = 32;     // columns of the A allocation: hi copy (X: 16, H1: 24 columns) | lo copy
// D[ACC] = A (tensor memory: lane = row, one TF32 element per column, KS slices of 8 columns; hi and lo copies)
//          * B (shared memory [k/4][n][4], NB n-rows per chunk), N = 32, 3xTF32
template <int KS>
__device__ __forceinline__ void gemm_k_tf32_16(uint32_t tmem_acc, uint32_t tmem_a, uint32_t b_hi16, uint32_t b_lo16) {
  static_assert(KS == 2 || KS == 3, "two or three 8-wide K slices");
  constexpr uint32_t id = tc::idesc_tf32(128, 32, false, false);
  const uint64_t dbh = tc::smem_desc16(b_hi16, NB * 16, 128), dbl = tc::smem_desc16(b_lo16, NB * 16, 128);
  const uint32_t ahi = tmem_a + C_AHI, alo = tmem_a + C_ALO;
#define FBSDEJ_TF32_SLICE(P0)                                                           \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %5, " P0 ";\n\t"                  \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %5, pt;\n\t"                      \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %5, pt;\n\t"
#define FBSDEJ_TF32_NEXT "add.u32 ah, ah, 8;\n\tadd.u32 al, al, 8;\n\tadd.s64 bh, bh, %6;\n\tadd.s64 bl, bl, %6;\n\t"
  if constexpr (KS == 2) {
    asm volatile(
        "{\n\t.reg .pred pf, pt;\n\t.reg .b32 ah, al;\n\t.reg .b64 bh, bl;\n\t"
        "setp.ne.u32 pf, 0, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
        "mov.b32 ah, %1;\n\tmov.b32 al, %2;\n\tmov.b64 bh, %3;\n\tmov.b64 bl, %4;\n\t"
        FBSDEJ_TF32_SLICE("pf") FBSDEJ_TF32_NEXT FBSDEJ_TF32_SLICE("pt") "}"
        :
        : "r"(tmem_acc), "r"(ahi), "r"(alo), "l"(dbh), "l"(dbl), "r"(id), "n"(2 * NB)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pf, pt;\n\t.reg .b32 ah, al;\n\t.reg .b64 bh, bl;\n\t"
        "setp.ne.u32 pf, 0, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
        "mov.b32 ah, %1;\n\tmov.b32 al, %2;\n\tmov.b64 bh, %3;\n\tmov.b64 bl, %4;\n\t"
        FBSDEJ_TF32_SLICE("pf") FBSDEJ_TF32_NEXT FBSDEJ_TF32_SLICE("pt") FBSDEJ_TF32_NEXT FBSDEJ_TF32_SLICE("pt") "}"
        :
        : "r"(tmem_acc), "r"(ahi), "r"(alo), "l"(dbh), "l"(dbl), "r"(id), "n"(2 * NB)
        : "memory");
  }
#undef FBSDEJ_TF32_SLICE
#undef FBSDEJ_TF32_NEXT
}
template <int KS>
__device__ __forceinline__ void gemm_k_tf32(uint32_t tmem_acc, uint32_t tmem_a, uint32_t b_hi, uint32_t b_lo) {
  gemm_k_tf32_16<KS>(tmem_acc, tmem_a, tc::addr16(b_hi), tc::addr16(b_lo));
}
// 8 consecutive features of this thread's row -> TF32 hi / lo columns of the A operand in tensor memory
__device__ __forceinline__ void store_tf32x8(uint32_t lane_base /* of the A allocation */, int c8, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float hi, lo;
    tc::split_tf32(v[q], hi, lo);
    h[q] = __float_as_uint(hi); l[q] = __float_as_uint(lo);
  }
  tc::tmem_st8(lane_base + C_AHI + 8 * c8, h);
  tc::tmem_st8(lane_base + C_ALO + 8 * c8, l);
}
// TMEM operand writes (+ the bias row in shared memory) -> visible to the MMA issued after the CTA barrier
__device__ __forceinline__ void publish_tmem() {
  tc::tmem_st_wait();
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
}
}  // namespace fwd

}  // namespace rtc
}  // namespace fbsdej

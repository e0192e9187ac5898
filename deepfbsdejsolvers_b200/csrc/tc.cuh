// sm_100a tensor-core plumbing used by the tcgen05 variants of the fused kernels: mbarriers, TMEM allocation,
// UMMA shared-memory / instruction descriptors (no-swizzle canonical layouts), tcgen05.mma kind::tf32, tcgen05.ld.
//
// Operand layouts (fp32 containers, TF32 math).  Every activation tile of tile_mlp.cuh is stored as float4 chunks
// [feature/4][row][4]; that single byte layout is
//   * a K-major canonical operand when features are the K dimension (forward / input-gradient GEMMs): 8-row x 16-byte
//     core matrices, leading-byte-offset (next 16-byte chunk along K) = 2048 B, stride-byte-offset (next 8 rows) = 128 B;
//   * an MN-major canonical operand when rows are the K dimension (weight-gradient GEMMs): the 8 rows of one K step
//     are contiguous (128 B), the next 4-feature chunk along M/N is 2048 B away.
// fp32 accuracy from TF32 tensor cores: operands are split x = hi + lo with hi = x with the 13 low mantissa bits
// cleared (exactly representable in TF32) and lo = x - hi (exact in fp32); D += A_hi B_hi + A_lo B_hi + A_hi B_lo.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace fbsdej {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: the warp sleeps in hardware instead of spinning
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP): `bytes` and both addresses are multiples of 16; completion is
// signalled on the mbarrier as a transaction count, so the issuing thread first arms it with expect_tx.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------------------------
// Called by ONE full warp.  ncols: power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols, bool last = true) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  if (last) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");   // after the CTA's last allocation
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
// shared-memory address in units of 16 bytes, as the 14-bit descriptor field wants it.  The mask matters: in a thread-block
// cluster the shared::cta window of a CTA does not start at 0, and the unmasked high bits would run into the LBO field.
__device__ __forceinline__ uint32_t addr16(uint32_t smem_addr) { return (smem_addr >> 4) & 0x3FFFu; }
// The same descriptor from the operand's shared-memory address IN UNITS OF 16 BYTES (addr16 = tc::addr16(address)): with addr16 =
// (loop-invariant base >> 4) + compile-time offset the low word is ONE add and the high word a constant - the byte-address form
// above costs a shift, a mask and an or per descriptor, twice per MMA, on the issuing thread's critical path.
__device__ __forceinline__ uint64_t smem_desc16(uint32_t addr16, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = addr16 + (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
// 32-bit instruction descriptor, kind::tf32, fp32 accumulate.  a_mn / b_mn: operand is MN-major.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with BF16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// One lane of a converged warp (elect.sync): the issuing-thread predicate ptxas recognises - a `lane == 0` test makes it wrap every
// tcgen05 instruction of the branch in an elect / retry loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
  return p != 0;
}
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in TMEM (lane = row; K-major: kind::tf32 one element per 32-bit column, kind::f16 two packed elements per
// column, the even element in the low half), B operand in shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all tcgen05 operations issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: lane (= row) of this thread, 8 consecutive columns -------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: lane (= row) of this thread, 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// hi/lo split for 3xTF32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
// bf16 hi/lo split (bf16x3: D += A_hi B_hi + A_lo B_hi + A_hi B_lo, ~2^-16 relative): returns the two bf16 bit patterns
__device__ __forceinline__ void split_bf16(float x, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = (uint32_t)__bfloat16_as_ushort(h);
  lo = (uint32_t)__bfloat16_as_ushort(l);
}
// Store 8 consecutive features of this thread's row as bf16 hi / lo chunks: tile layout [feature/8][row][8 bf16].
// hi = the upper 16 bits of the fp32 value (truncation, exactly representable), lo = bf16_rn(x - hi): two values are
// packed per PRMT / cvt.rn.bf16x2 instruction.
__device__ __forceinline__ void store_bf16x8(uint4* __restrict__ hi_tile, uint4* __restrict__ lo_tile, int chunk, int row,
                                             const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t b0 = __float_as_uint(v[2 * i]), b1 = __float_as_uint(v[2 * i + 1]);
    h[i] = __byte_perm(b0, b1, 0x7632);                                   // {hi16(b0), hi16(b1)}
    const float r0 = v[2 * i] - __uint_as_float(b0 & 0xFFFF0000u), r1 = v[2 * i + 1] - __uint_as_float(b1 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(r1), "f"(r0));   // upper half <- first source
  }
  hi_tile[chunk * 128 + row] = make_uint4(h[0], h[1], h[2], h[3]);
#if !(defined(FBSDEJ_ABLATE) && FBSDEJ_ABLATE == 14)
  lo_tile[chunk * 128 + row] = make_uint4(l[0], l[1], l[2], l[3]);
#endif
}

}  // namespace tc
}  // namespace fbsdej

// Common device/host helpers for the fbsdej sm_100a library.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace fbsdej {

constexpr int kThreads = 128;        // threads per CTA of the fused path kernels (= rows of one MLP tile)
constexpr int kHeader = 4;           // out vector header: [loss, loss_a, loss_b, aux]

// ---- error plumbing (C-ABI returns codes; message kept thread-local) -------------------------
void set_error(const std::string& msg);
#define FB_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ::fbsdej::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));               \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)
#define FB_REQUIRE(cond, msg)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::fbsdej::set_error(std::string(msg));                                                  \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) ---------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline uint4 rand4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
      uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
      uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
      uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
      uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
      uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// stream ids of the counter's 4th word
// Primary streams: the per-path increments, the compensator samples (word 0 = sample index), the MFG increments.  Every
// secondary draw (a retry of a rejection sampler, the second block of a rare Merton jump) has its OWN named stream.
enum : uint32_t {
  STREAM_PATH = 0, STREAM_JMC = 1,
  STREAM_PATH_RARE_JUMP = 2, STREAM_JMC_RARE_JUMP = 3,       // Merton: count >= 2 (sim_device.cuh: jump_size_rare)
  STREAM_PATH_GAMMA_RETRY = 4, STREAM_JMC_GAMMA_RETRY = 5,   // VG: Marsaglia-Tsang rejections (sim_vg_kernel)
  STREAM_MFG = 6, STREAM_MFG_POISSON_RETRY = 7               // MFG: PTRS rejections (poisson_any)
};
__host__ __device__ constexpr uint32_t rare_jump_stream(uint32_t s) { return s == STREAM_PATH ? STREAM_PATH_RARE_JUMP : STREAM_JMC_RARE_JUMP; }
__host__ __device__ constexpr uint32_t gamma_retry_stream(uint32_t s) { return s == STREAM_PATH ? STREAM_PATH_GAMMA_RETRY : STREAM_JMC_GAMMA_RETRY; }
static_assert(rare_jump_stream(STREAM_PATH) != rare_jump_stream(STREAM_JMC) && gamma_retry_stream(STREAM_PATH) != gamma_retry_stream(STREAM_JMC) &&
              STREAM_MFG_POISSON_RETRY != STREAM_MFG && STREAM_MFG > STREAM_JMC_GAMMA_RETRY, "Philox stream ids must be distinct");

__device__ __forceinline__ float u01_open(uint32_t x) {  // (0,1]
  return (float)((x >> 8) + 1u) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float u01_half(uint32_t x) {  // [0,1)
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}
// Box-Muller: two N(0,1) from two 32-bit words.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float r = sqrtf(-2.0f * __logf(u01_open(a)));
  float s, c;
  __sincosf(6.283185307179586f * u01_half(b), &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// ---- reductions -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over the G consecutive threads of a path group (G power of two <= 32): every lane gets the sum.
__device__ __forceinline__ float group_sum_shfl(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum (kThreads threads), result valid in every thread. `red` = 4 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__host__ __device__ constexpr int round_up4(int x) { return (x + 3) & ~3; }

// Standard normal CDF the way tfp does it: 0.5*erfc(-x/sqrt(2)).
__device__ __forceinline__ float ncdf(float x) { return 0.5f * erfcf(-x * 0.7071067811865476f); }

}  // namespace fbsdej

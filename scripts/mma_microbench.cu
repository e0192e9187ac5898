// Micro-benchmark behind DESIGN.md section 3.1 ("what a small tcgen05.mma costs"): one thread issues a train of NM MMAs,
// commits, and the CTA waits on the mbarrier; cycles / MMA are reported for
//   kind      0: bf16 SS, K-major operands (layer GEMMs of the adjoint)     1: bf16 SS, MN-major (weight-gradient GEMMs)
//             2: tf32 TS (A operand in tensor memory, forward sweep)        3: bf16 TS
//   N         accumulator width
//   nacc      number of distinct accumulators the train rotates over (1 = every MMA depends on the previous one)
//   M         128 or 64
// with one CTA on the device and with 4 CTAs on every SM (the training kernels' residency).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/mma_microbench scripts/mma_microbench.cu
//        add -DMB_ELECT to pick the issuing lane with elect.sync (tc::elect_one) instead of a lane test: ptxas then emits the
//        tcgen05 instructions without an elect / retry loop around each (profiles/r2_mma_microbench_elect.csv)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../deepfbsdejsolvers_b200/csrc/tc.cuh"

using namespace fbsdej;

constexpr int NM = 64;
#ifdef MB_ELECT
#define MB_ISSUER0 ((threadIdx.x >> 5) == 0 && tc::elect_one())
#define MB_ISSUERW(w, NT) ((w) < (NT) && tc::elect_one())
#else
#define MB_ISSUER0 (threadIdx.x == 0)
#define MB_ISSUERW(w, NT) ((threadIdx.x & 31) == 0 && (w) < (NT))
#endif

template <int KIND, int N, int NACC, int M>
__global__ void __launch_bounds__(128) bench(int ncols, long long* out) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 16384; i += 128) sm[i] = 0.0f;
  if (threadIdx.x < 32) tc::tmem_alloc(&tslot, ncols);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tslot;
  const uint32_t sb = tc::smem_u32(sm);
  long long t0 = 0, t1 = 0;
  constexpr uint32_t acc0 = (KIND >= 2) ? 32 : 0;      // A operand (TS kinds) lives in the first 32 columns
  constexpr uint32_t idk = KIND == 2 ? tc::idesc_tf32(M, N, false, false) : tc::idesc_bf16(M, N, KIND == 1, KIND == 1);
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    if (MB_ISSUER0) {
      tc::tc_fence_after();
      uint64_t da[8], db[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        da[i] = KIND == 1 ? tc::smem_desc(sb + i * 256, 128, 2048) : tc::smem_desc(sb + (i & 1) * 4096, 2048, 128);
        db[i] = KIND == 1 ? tc::smem_desc(sb + 32768 + i * 256, 128, 2048) : tc::smem_desc(sb + 32768, N * 16, 128);
      }
      t0 = clock64();
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        const uint32_t d = tm + acc0 + (uint32_t)((i % NACC) * N);
        const uint32_t acc = i >= NACC ? 1u : 0u;
        if (KIND <= 1) tc::mma_bf16(d, da[i & 7], db[i & 7], idk, acc);
        else if (KIND == 2) tc::mma_tf32_ts(d, tm + (i & 3) * 8, db[0], idk, acc);
        else tc::mma_bf16_ts(d, tm + (i & 3) * 8, db[0], idk, acc);
      }
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, rep & 1);
    tc::tc_fence_after();
    if (threadIdx.x == 0) t1 = clock64();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tm, ncols);
}

static const char* names[4] = {"bf16 SS K-major", "bf16 SS MN-major", "tf32 TS", "bf16 TS"};
template <int KIND, int N, int NACC, int M>
static void run(int grid, long long* d) {
  constexpr int need = (KIND >= 2 ? 32 : 0) + NACC * N;
  int ncols = 32;
  while (ncols < need) ncols *= 2;
  if (ncols > 512 || (grid > 1 && ncols > 128)) return;     // 3 CTAs x <= 128 columns per SM
  const int smem = 65536 + 2048;
  cudaFuncSetAttribute(bench<KIND, N, NACC, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<KIND, N, NACC, M><<<grid, 128, smem>>>(ncols, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s N=%d M=%d nacc=%d: %s\n", names[KIND], N, M, NACC, cudaGetErrorString(e)); exit(1); }
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s,%d,%d,%d,%d,%.1f\n", names[KIND], N, M, NACC, grid, (double)h / NM);
}
template <int KIND, int M>
static void run_kind(int grid, long long* d) {
  run<KIND, 16, 1, M>(grid, d); run<KIND, 16, 2, M>(grid, d); run<KIND, 16, 4, M>(grid, d);
  run<KIND, 32, 1, M>(grid, d); run<KIND, 32, 2, M>(grid, d); run<KIND, 32, 4, M>(grid, d);
  run<KIND, 48, 1, M>(grid, d); run<KIND, 48, 2, M>(grid, d);
  run<KIND, 64, 1, M>(grid, d); run<KIND, 64, 2, M>(grid, d);
  run<KIND, 96, 1, M>(grid, d); run<KIND, 128, 1, M>(grid, d);
  if (KIND != 1) run<KIND, 256, 1, M>(grid, d);
}

// NT threads (lane 0 of warps 0..NT-1) of ONE CTA issue NM MMAs each, every thread into its own accumulator and with its own
// commit on a shared mbarrier (count NT): is the ~46-cycle issue floor per thread or per CTA?
template <int KIND, int N, int NT>
__global__ void __launch_bounds__(128) bench_mt(long long* out) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 16384; i += 128) sm[i] = 0.0f;
  if (threadIdx.x < 32) tc::tmem_alloc(&tslot, 512);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, NT); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tslot;
  const uint32_t sb = tc::smem_u32(sm);
  long long t0 = 0, t1 = 0;
  constexpr uint32_t idk = KIND == 2 ? tc::idesc_tf32(128, N, false, false) : tc::idesc_bf16(128, N, false, false);
  const int w = threadIdx.x >> 5;
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    t0 = clock64();
    if (MB_ISSUERW(w, NT)) {
      tc::tc_fence_after();
      const uint64_t da = tc::smem_desc(sb + w * 4096, 2048, 128), db = tc::smem_desc(sb + 32768, N * 16, 128);
      const uint32_t d = tm + 32 + (uint32_t)(w * N);
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        if (KIND == 0) tc::mma_bf16(d, da, db, idk, i ? 1u : 0u);
        else tc::mma_tf32_ts(d, tm + (i & 3) * 8, db, idk, i ? 1u : 0u);
      }
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, rep & 1);
    tc::tc_fence_after();
    t1 = clock64();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tm, 512);
}
template <int KIND, int N, int NT>
static void run_mt(long long* d) {
  const int smem = 65536 + 2048;
  cudaFuncSetAttribute(bench_mt<KIND, N, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench_mt<KIND, N, NT><<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mt %s N=%d NT=%d: %s\n", names[KIND], N, NT, cudaGetErrorString(e)); exit(1); }
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("multi-thread %s,%d,128,%d threads,1,%.1f\n", names[KIND], N, NT, (double)h / (NM * NT));
}

// One "layer evaluation" round trip as the fused kernels do it: every thread writes its row of the A operand to tensor memory
// (tcgen05.st), wait::st + fences + CTA barrier, ONE thread issues NMMA tf32 MMAs (A from TMEM) and commits, all threads wait on
// the mbarrier and read 24 accumulator columns back.  Cycles per round trip = the latency floor of one network layer of a
// lone CTA (configs 1 / 2 / 4 walk N time steps of 2 (forward) + 4 (adjoint) such round trips on a handful of CTAs).
template <int NMMA>
__global__ void __launch_bounds__(128) bench_roundtrip(long long* out, float* sink) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 16384; i += 128) sm[i] = 0.0f;
  if (threadIdx.x < 32) tc::tmem_alloc(&tslot, 128);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tslot, lane = tm + ((uint32_t)(threadIdx.x & ~31) << 16);
  const uint32_t sb = tc::smem_u32(sm);
  constexpr uint32_t idk = tc::idesc_tf32(128, 32, false, false);
  constexpr int R = 256;
  float acc = 0.0f;
  uint32_t phase = 0;
  const long long t0 = clock64();
  for (int r = 0; r < R; ++r) {
    uint32_t v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __float_as_uint(acc + (float)q);
    tc::tmem_st8(lane + 64, v);
    tc::tmem_st_wait();
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (MB_ISSUER0) {
      tc::tc_fence_after();
      const uint64_t db = tc::smem_desc(sb + 32768, 32 * 16, 128);
#pragma unroll
      for (int i = 0; i < NMMA; ++i) tc::mma_tf32_ts(tm, tm + 64, db, idk, i ? 1u : 0u);
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, phase); phase ^= 1;
    tc::tc_fence_after();
    float a8[8], b8[8], c8[8];
    tc::tmem_ld8(lane, a8); tc::tmem_ld8(lane + 8, b8); tc::tmem_ld8(lane + 16, c8);
    tc::tmem_ld_wait();
    acc = a8[0] + b8[1] + c8[2];
    tc::tc_fence_before();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / R;
  sink[blockIdx.x * 128 + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tm, 128);
}
template <int NMMA>
static void run_rt(long long* d, float* sink) {
  const int smem = 65536 + 2048;
  cudaFuncSetAttribute(bench_roundtrip<NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148 * 3}) {
    bench_roundtrip<NMMA><<<grid, 128, smem>>>(d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("roundtrip NMMA=%d: %s\n", NMMA, cudaGetErrorString(e)); exit(1); }
    long long h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("round trip (st + barrier + %d tf32 TS MMAs + commit + wait + ld),32,128,1,%d,%lld\n", NMMA, grid, h);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  printf("kind,N,M,nacc,grid,cycles_per_mma\n");
  float* sink;
  cudaMalloc(&sink, 148 * 3 * 128 * 4);
  run_rt<1>(d, sink); run_rt<6>(d, sink); run_rt<9>(d, sink);
  run_mt<0, 32, 1>(d); run_mt<0, 32, 2>(d); run_mt<0, 32, 4>(d); run_mt<2, 32, 1>(d); run_mt<2, 32, 2>(d); run_mt<2, 32, 4>(d);
  run_mt<0, 64, 2>(d); run_mt<2, 64, 2>(d);
  for (int grid : {1, 148 * 3}) {
    run_kind<0, 128>(grid, d); run_kind<1, 128>(grid, d); run_kind<2, 128>(grid, d); run_kind<3, 128>(grid, d);
  }
  for (int grid : {1, 148 * 3}) { run_kind<0, 64>(grid, d); run_kind<2, 64>(grid, d); }
  return 0;
}

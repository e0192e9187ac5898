// Increment simulation: counter-based Philox4x32-10, shard-invariant (counter word 0 = GLOBAL path id).
//
// Replaces the stateful TF draws of the reference:
//   dW = sqrt(dt)*tf.random.normal            SolversJumpDiff.py:30-31, MFGSolvers.py:35-36
//   Merton jumps: Poisson(lam dt), N(0,1)     pricingModels.py:57-61
//   VG jumps: Gamma(dt/kappa, rate 1/kappa)   pricingModels.py:188-191
//   MFG Cox counts: Poisson(lam(hQ_i) dt)     MFGModel.py:47-54 (intensity frozen at the step start)
// Output layout: time-major component planes [N][d][B]; each thread owns 4 consecutive paths of one
// (step, component pair) and issues 128-bit stores.
#include "sim.cuh"
#include "sim_device.cuh"

namespace fbsdej {

// Poisson(mean) by table inversion: thr[k] = floor(CDF(k) * 2^32); count = #{k : u >= thr[k]}.
// One (path, asset pair, step) cell = ONE Philox call: words 0/1 -> the two Brownian normals (Box-Muller), words 2/3 ->
// the two Poisson counts by table inversion.  A count of 1 (the only frequent non-zero case: P = lam dt e^{-lam dt})
// takes its jump size from the SAME uniform: conditional on thr[0] <= u < thr[1], (u - thr[0]) / (thr[1] - thr[0]) is
// uniform with ~28 bits, mapped through the inverse normal CDF (branch-free polynomial, evaluated for every draw).
// Only counts >= 2 (P ~ (lam dt)^2 / 2) and far-tail sizes go through jump_size_rare (a second Philox block on its own stream, common.cuh).
// Work split: a work unit = (step, asset pair, chunk of 256 groups of 4 consecutive paths); a persistent grid (4 CTAs per
// SM) strides over the units, so the per-CTA set-up is paid once and the only integer divisions are per unit and uniform.
// 128-bit stores.
__global__ void __launch_bounds__(256, 4) sim_merton_kernel(const SimMertonArgs a) {
  __shared__ uint32_t sthr[64];
  if (threadIdx.x < 64) sthr[threadIdx.x] = threadIdx.x < a.npois ? a.pois_thr[threadIdx.x] : 0xffffffffu;
  __syncthreads();
  const int B4 = (a.B + 3) / 4, KP = (a.D + 1) / 2;
  const uint32_t iter = a.iter_ptr ? *a.iter_ptr : a.iteration;
  const uint32_t t0 = sthr[0], t1 = sthr[1];
  const float inv_w1 = t1 > t0 ? 1.0f / (float)(t1 - t0) : 0.0f;
  const bool vec = (a.B % 4 == 0);
  const int nchunk = (B4 + 255) / 256;
  const long long nunits = (long long)a.N * KP * nchunk;
  for (long long unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int cell = (int)(unit / nchunk), chunk = (int)(unit - (long long)cell * nchunk);   // (uniform per block)
    const int i = cell / KP, kp = cell - i * KP;
    const int k0 = 2 * kp, k1 = 2 * kp + 1;
    const uint32_t c1 = ((uint32_t)i << 8) | (uint32_t)kp;
    const size_t o0 = ((size_t)i * a.D + k0) * a.B, o1 = ((size_t)i * a.D + k1) * a.B;
    const int bq = chunk * 256 + threadIdx.x;
    if (bq < B4) {
      const int b0 = bq * 4;
      float w0[4], w1[4], j0[4], j1[4];
      uint32_t rare = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                        // branch-free: the four Philox chains interleave
        const uint32_t gid = a.path_offset + (uint32_t)(b0 + q);
        const uint4 r = Philox::rand4(gid, c1, iter, a.stream, a.seed_lo, a.seed_hi);
        box_muller_fast(r.x, r.y, a.sqdt, w0[q], w1[q]);
        uint32_t rq;
        jump_sizes_fast(r.z, r.w, t0, t1, inv_w1, a.muJ, a.sigJ, j0[q], j1[q], rq);
        rare |= rq << (2 * q);
      }
      if (rare) {                                          // ~0.4 % of the cells: redo the block, exact path
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if ((rare >> (2 * q)) & 3u) {
            const uint32_t gid = a.path_offset + (uint32_t)(b0 + q);
            const uint4 r = Philox::rand4(gid, c1, iter, a.stream, a.seed_lo, a.seed_hi);
            if ((rare >> (2 * q)) & 1u) j0[q] = jump_size_rare(r.z, t0, t1, inv_w1, sthr, a.npois, a.muJ, a.sigJ, gid, c1, iter, a.stream, a.seed_lo, a.seed_hi, 0);
            if ((rare >> (2 * q)) & 2u) j1[q] = jump_size_rare(r.w, t0, t1, inv_w1, sthr, a.npois, a.muJ, a.sigJ, gid, c1, iter, a.stream, a.seed_lo, a.seed_hi, 1);
          }
        }
      }
      if (vec) {
        if (a.dW) st4(a.dW + o0 + b0, make_float4(w0[0], w0[1], w0[2], w0[3]));
        st4(a.J + o0 + b0, make_float4(j0[0], j0[1], j0[2], j0[3]));
        if (k1 < a.D) {
          if (a.dW) st4(a.dW + o1 + b0, make_float4(w1[0], w1[1], w1[2], w1[3]));
          st4(a.J + o1 + b0, make_float4(j1[0], j1[1], j1[2], j1[3]));
        }
      } else {
        for (int q = 0; q < 4 && b0 + q < a.B; ++q) {
          if (a.dW) a.dW[o0 + b0 + q] = w0[q];
          a.J[o0 + b0 + q] = j0[q];
          if (k1 < a.D) {
            if (a.dW) a.dW[o1 + b0 + q] = w1[q];
            a.J[o1 + b0 + q] = j1[q];
          }
        }
      }
    }
  }
}

// Gamma(shape, 1) by Marsaglia-Tsang (shape >= 1) with the U^{1/shape} boost below 1.
// Rejection => variable Philox consumption: attempts walk a per-(path, step) sub-counter in word 1's high bits.
__global__ void __launch_bounds__(256) sim_vg_kernel(const SimVGArgs a) {
  const size_t total = (size_t)a.N * a.B;
  const uint32_t iter = a.iter_ptr ? *a.iter_ptr : a.iteration;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(t % a.B), i = (int)(t / a.B);
    const uint32_t gid = a.path_offset + (uint32_t)b;
    const uint4 r = Philox::rand4(gid, (uint32_t)i, iter, a.stream, a.seed_lo, a.seed_hi);
    float eps, x;
    box_muller(r.x, r.y, eps, x);
    const float ub = u01_open(r.z);
    float ua = u01_open(r.w);
    const bool boost = a.shape < 1.0f;
    const float al = boost ? a.shape + 1.0f : a.shape;
    const float d = al - (1.0f / 3.0f), c = rsqrtf(9.0f * d);
    float gam = d;
    float x2 = 0.0f, ua2 = 0.0f;
    bool have2 = false;
    for (uint32_t att = 0; att < 64u; ++att) {
      float v = 1.0f + c * x;
      if (v > 0.0f) {
        v = v * v * v;
        if (__logf(ua) < 0.5f * x * x + d - d * v + d * __logf(v)) { gam = d * v; break; }
      }
      if (have2) { x = x2; ua = ua2; have2 = false; }
      else {
        const uint4 s = Philox::rand4(gid, (uint32_t)i | ((att / 2u + 1u) << 24), iter, gamma_retry_stream(a.stream), a.seed_lo, a.seed_hi);
        box_muller(s.x, s.y, x, x2);
        ua = u01_open(s.z); ua2 = u01_open(s.w);
        have2 = true;
      }
    }
    if (boost) gam *= __powf(ub, 1.0f / a.shape);
    gam *= a.scale;                                                  // rate 1/kappa -> scale kappa
    a.J[(size_t)i * a.B + b] = a.theta * gam + a.sigJ * sqrtf(gam) * eps;   // pricingModels.py:191
  }
}

// Poisson(mean) for any mean: sequential inversion below 10, PTRS transformed rejection (Hormann 1993) above.
__device__ float poisson_any(float mean, uint32_t u32a, uint32_t u32b, uint32_t gid, uint32_t c1, uint32_t iter,
                             uint32_t stream, uint32_t k0, uint32_t k1) {
  if (!(mean > 0.0f)) return 0.0f;
  if (mean < 10.0f) {
    const float u = ((float)(u32a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float p = __expf(-mean), s = p;
    int k = 0;
    while (u > s && k < 200) { ++k; p *= mean / (float)k; s += p; }
    return (float)k;
  }
  const float slam = sqrtf(mean), loglam = __logf(mean);
  const float b = 0.931f + 2.53f * slam, aa = -0.059f + 0.02483f * b;
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.0f);
  uint32_t ua = u32a, ub = u32b;
  for (uint32_t att = 0; att < 64u; ++att) {
    const float U = u01_half(ua) - 0.5f, V = u01_open(ub);
    const float us = 0.5f - fabsf(U);
    const float kf = floorf((2.0f * aa / us + b) * U + mean + 0.43f);
    if (us >= 0.07f && V <= vr) return kf;
    if (!(kf < 0.0f || (us < 0.013f && V > us))) {
      if (__logf(V) + __logf(invalpha) - __logf(aa / (us * us) + b) <= -mean + kf * loglam - lgammaf(kf + 1.0f)) return kf;
    }
    const uint4 s = Philox::rand4(gid, c1 | ((att + 1u) << 24), iter, STREAM_MFG_POISSON_RETRY, k0, k1);
    ua = s.x; ub = s.y;
  }
  return floorf(mean + 0.5f);
}

// MFG: thread per path walks the exogenous hQ recursion to get the Cox intensity of every step.
__global__ void __launch_bounds__(256) sim_mfg_kernel(const SimMFGArgs a) {
  const uint32_t iter = a.iter_ptr ? *a.iter_ptr : a.iteration;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.B; b += gridDim.x * blockDim.x) {
    const uint32_t gid = a.path_offset + (uint32_t)b;
    float hQ = a.q0;
    for (int i = 0; i < a.N; ++i) {
      const uint4 r = Philox::rand4(gid, (uint32_t)i, iter, a.stream, a.seed_lo, a.seed_hi);
      float n0, n1;
      box_muller(r.x, r.y, n0, n1);
      const float dW0 = a.sqdt * n0, dW = a.sqdt * n1;
      const float lam = a.stochastic ? a.beta * (expf(a.alpha * hQ) - 1.0f) : a.jumpFactor;   // MFGModel.py:49-52
      const float dN = poisson_any(lam * a.dt, r.z, r.w, gid, (uint32_t)i, iter, a.stream, a.seed_lo, a.seed_hi);
      a.dW0[(size_t)i * a.B + b] = dW0;
      a.dW[(size_t)i * a.B + b] = dW;
      a.dN[(size_t)i * a.B + b] = dN;
      hQ = hQ + a.coeffOU * (a.qaver[i + 1] - hQ) * a.dt + a.sig0 * dW0;                      // MFGModel.py:70
    }
  }
}

// Stable compaction of the compensator samples of each step: non-zero samples first, count of all-zero ones.
// (Merton: ~94 % of the 5000 samples are exactly 0 at lam*dt = 0.06; their mean contribution is n0*G(i,X,0)/M.)
constexpr int kCompactThreads = 1024;   // one block per step: fewer (barrier-separated) rounds over the M samples
__global__ void __launch_bounds__(kCompactThreads) compact_jmc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                     int* __restrict__ nnz, int* __restrict__ n0, int D, int M,
                                                                     int dedup) {
  constexpr int NW = kCompactThreads / 32;
  __shared__ int swarp[NW];
  __shared__ int sbase;
  const int i = blockIdx.x;
  const float* s = src + (size_t)i * D * M;
  float* d = dst + (size_t)i * D * M;
  if (threadIdx.x == 0) sbase = 0;
  __syncthreads();
  for (int m0 = 0; m0 < M; m0 += kCompactThreads) {
    const int m = m0 + threadIdx.x;
    bool nz = false;
    if (m < M) {
      if (!dedup) nz = true;
      else for (int k = 0; k < D; ++k) nz = nz || (s[(size_t)k * M + m] != 0.0f);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, nz);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) swarp[wid] = __popc(bal);
    __syncthreads();
    int off = sbase;
    for (int w = 0; w < wid; ++w) off += swarp[w];
    off += __popc(bal & ((1u << lane) - 1u));
    if (nz) for (int k = 0; k < D; ++k) d[(size_t)k * M + off] = s[(size_t)k * M + m];
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < NW; ++w) tot += swarp[w]; sbase += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { nnz[i] = sbase; n0[i] = M - sbase; }
}

static inline int sim_grid(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

int launch_sim_merton(const SimMertonArgs& a, cudaStream_t st) {
  const long long nunits = (long long)a.N * ((a.D + 1) / 2) * (((a.B + 3) / 4 + 255) / 256);
  const int grid = (int)(nunits < 148 * 4 ? (nunits < 1 ? 1 : nunits) : 148 * 4);
  sim_merton_kernel<<<grid, 256, 0, st>>>(a);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_sim_vg(const SimVGArgs& a, cudaStream_t st) {
  sim_vg_kernel<<<sim_grid((size_t)a.N * a.B, 256), 256, 0, st>>>(a);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_sim_mfg(const SimMFGArgs& a, cudaStream_t st) {
  sim_mfg_kernel<<<sim_grid((size_t)a.B, 256), 256, 0, st>>>(a);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_compact_jmc(const float* src, float* dst, int* nnz, int* n0, int N, int D, int M, int dedup, cudaStream_t st) {
  compact_jmc_kernel<<<N, kCompactThreads, 0, st>>>(src, dst, nnz, n0, D, M, dedup);
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fbsdej

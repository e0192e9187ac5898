#include "util.cuh"

namespace fbsdej {

// out[0..3] = sum over CTAs of the loss partials; out[4+e] = sum over CTAs of gradient partials (fixed order:
// deterministic for a given grid).  A block of 256 threads owns 32 consecutive output elements; its 8 warps split the
// partial rows (row c goes to warp c mod 8, 8 independent loads in flight per thread), then one fixed-order sum over warps.
// With theta != NULL the kernel also finishes the training step: Keras-form Adam on its elements and, in the block that
// finishes last, the step / iteration counters and the loss record (replaces four more launches per step).
struct FinishArgs {
  float* theta; float* m; float* v; const float* mask;
  float lr, b1, b2, eps;
  int* t_dev; uint32_t* iter_dev; float* loss_dst; uint32_t* step_ctr; unsigned int* done_ctr;
  XchgArgs x;                    // x.world > 1: sum the vector over the ranks before the update
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ lpart, int nparts_l,
                                                              const float* __restrict__ gpart, int nparts, int P,
                                                              float* __restrict__ out, int with_grad, const FinishArgs f) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + tx;
  const int n = kHeader + (with_grad ? P : 0);
  float s = 0.0f;
  if (e < kHeader) {
    for (int c = ty; c < nparts_l; c += 8) s += lpart[c * 4 + e];
  } else if (e < n) {
    const float* __restrict__ g = gpart + (e - kHeader);
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, a4 = 0.0f, a5 = 0.0f, a6 = 0.0f, a7 = 0.0f;
    int c = ty;
    for (; c + 56 < nparts; c += 64) {
      a0 += g[(size_t)c * P]; a1 += g[(size_t)(c + 8) * P]; a2 += g[(size_t)(c + 16) * P]; a3 += g[(size_t)(c + 24) * P];
      a4 += g[(size_t)(c + 32) * P]; a5 += g[(size_t)(c + 40) * P]; a6 += g[(size_t)(c + 48) * P]; a7 += g[(size_t)(c + 56) * P];
    }
    for (; c < nparts; c += 8) a0 += g[(size_t)c * P];
    s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  }
  red[ty][tx] = s;
  __syncthreads();
  float t = 0.0f;
  bool late = false;
  if (ty == 0) {
    t = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][tx];
    if (f.x.world > 1) {
      // Data-parallel step: this block's 32 elements go to every rank's slot of this rank (peer memory over NVLink, or the
      // same device in the single-process tests), the block's flag on every rank is raised to the step stamp, and once the
      // W flags of this block have arrived here the W rows are added in rank order - the same order, hence the same bits,
      // on every rank.  Two data slots alternate: a peer can be at most one step ahead (it needs this rank's next stamp).
      const XchgArgs& x = f.x;
      const uint32_t stamp = *x.xctr + 1u;
      const size_t slot = (size_t)(stamp & 1u) * x.world;
      for (int r = 0; r < x.world; ++r) x.peer_data[r][(slot + x.rank) * x.nstride + e] = t;
      __threadfence_system();
      __syncwarp();
      if (tx < x.world) st_release_sys(x.peer_flags[tx] + (size_t)x.rank * x.nblk + blockIdx.x, stamp);
      if (tx < x.world) {
        // a peer that never arrives (a rank died, or the ranks disagree on the number of steps) must not hang the device:
        // after x.timeout_ns the step is VOID - no parameter, Adam-slot or counter update on this rank - and the failed
        // stamp is written into the error word of every rank's buffer; fbsdej_solver_dp_check reports it to the host
        const uint32_t* fl = x.peer_flags[x.rank] + (size_t)tx * x.nblk + blockIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(fl) - stamp) < 0) {
          if (globaltimer_ns() - t0 > x.timeout_ns) { late = true; break; }
        }
      }
      late = __any_sync(0xffffffffu, late);
      if (late && tx < x.world) atomicCAS_system(x.peer_flags[tx] + (size_t)x.world * x.nblk + 1, 0u, stamp);
      // an error raised by ANY block of ANY rank voids the step here too (best effort: a rank that already passed this point
      // has updated; the host-side check still fails on every rank)
      late = late || ld_acquire_sys(x.xctr + 1) != 0u;
      const float* mine = x.peer_data[x.rank];
      t = 0.0f;
      for (int r = 0; r < x.world; ++r) t += __ldcv(mine + (slot + r) * x.nstride + e);
      if (late) t = __int_as_float(0x7fc00000);       // the loss / gradient record of a void step reads NaN
    }
  }
  late = __shfl_sync(0xffffffffu, late ? 1 : 0, 0) != 0;
  if (ty == 0 && e < n) {
    out[e] = t;
    if (f.theta && e >= kHeader && !late) {           // oracle/adam.py, SURVEY fact 9
      const int i = e - kHeader;
      if (!(f.mask && f.mask[i] == 0.0f)) {
        const int step = *f.t_dev + 1;
        const float b1p = powf(f.b1, (float)step), b2p = powf(f.b2, (float)step);
        const float alpha = f.lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
        const float mi = f.m[i] + (t - f.m[i]) * (1.0f - f.b1);
        const float vi = f.v[i] + (t * t - f.v[i]) * (1.0f - f.b2);
        f.m[i] = mi; f.v[i] = vi;
        f.theta[i] -= alpha * mi / (sqrtf(vi) + f.eps);
      }
    }
  }
  if (f.theta) {                                      // the block that finishes last advances the counters
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(f.done_ctr, 1u);
      if (prev == gridDim.x - 1) {
        __threadfence();
        const bool failed = f.x.world > 1 && ld_acquire_sys(f.x.xctr + 1) != 0u;   // void step: counters stay
        if (!failed) {
          *f.t_dev += 1;
          *f.iter_dev += 1u;
        }
        if (f.loss_dst) f.loss_dst[*f.step_ctr] = __ldcg(out);
        *f.step_ctr += 1u;
        if (f.x.world > 1) *f.x.xctr += 1u;
        *f.done_ctr = 0u;
      }
    }
  }
}

// Keras OptimizerV2 Adam (TF ResourceApplyAdam): see oracle/adam.py and SURVEY fact 9.
__global__ void adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ grad, const float* __restrict__ mask, int n, float lr, float b1,
                            float b2, float eps, const int* __restrict__ t_dev) {
  const int t = *t_dev + 1;
  const float b1p = powf(b1, (float)t), b2p = powf(b2, (float)t);
  const float alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (mask && mask[i] == 0.0f) continue;
    const float g = grad[i];
    const float mi = m[i] + (g - m[i]) * (1.0f - b1);
    const float vi = v[i] + (g * g - v[i]) * (1.0f - b2);
    m[i] = mi; v[i] = vi;
    theta[i] -= alpha * mi / (sqrtf(vi) + eps);
  }
}
__global__ void bump_i32_kernel(int* p) { *p += 1; }
__global__ void bump_u32_kernel(uint32_t* p) { *p += 1u; }
__global__ void copy_loss_kernel(const float* out, float* dst, uint32_t* ctr) {
  if (dst) dst[*ctr] = out[0];
  *ctr += 1u;
}

// Generic row-wise MLP (runtime dims, H <= 64): Net.call of the reference for Y0 reports and drop-in __call__.
__global__ void net_forward_kernel(const float* __restrict__ th, int nin, int H, int L, int nout, int act,
                                   const float* __restrict__ x, int rows, float* __restrict__ y) {
  extern __shared__ float sw[];
  const int np = nin * H + H + (L - 1) * (H * H + H) + H * nout + nout;
  for (int i = threadIdx.x; i < np; i += blockDim.x) sw[i] = th[i];
  __syncthreads();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    float h[64], h2[64];
    const float* w = sw;
    for (int j = 0; j < H; ++j) {
      float acc = w[nin * H + j];
      for (int k = 0; k < nin; ++k) acc = fmaf(x[(size_t)r * nin + k], w[k * H + j], acc);
      h[j] = act == 0 ? tanhf(acc) : fmaxf(acc, 0.0f);
    }
    w += nin * H + H;
    for (int l = 1; l < L; ++l) {
      for (int j = 0; j < H; ++j) {
        float acc = w[H * H + j];
        for (int k = 0; k < H; ++k) acc = fmaf(h[k], w[k * H + j], acc);
        h2[j] = act == 0 ? tanhf(acc) : fmaxf(acc, 0.0f);
      }
      for (int j = 0; j < H; ++j) h[j] = h2[j];
      w += H * H + H;
    }
    for (int j = 0; j < nout; ++j) {
      float acc = w[H * nout + j];
      for (int k = 0; k < H; ++k) acc = fmaf(h[k], w[k * nout + j], acc);
      y[(size_t)r * nout + j] = acc;
    }
  }
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int B, int d, int to_ndb) {
  const size_t total = (size_t)N * B * d;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    // t indexes the destination
    if (to_ndb) {
      const int b = (int)(t % B), k = (int)((t / B) % d), i = (int)(t / ((size_t)B * d));
      dst[t] = src[((size_t)i * B + b) * d + k];
    } else {
      const int k = (int)(t % d), b = (int)((t / d) % B), i = (int)(t / ((size_t)B * d));
      dst[t] = src[((size_t)i * d + k) * B + b];
    }
  }
}

__global__ void scatter_kernel(float* __restrict__ dst, const uint32_t* __restrict__ idx, const float* __restrict__ val, int nnz,
                               size_t n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += gridDim.x * blockDim.x) {
    const uint32_t k = idx[i];
    if (k < n) dst[k] = val[i];
  }
}
// dst[0 .. n) = 0, then dst[idx[i]] = val[i]
int launch_scatter(float* dst, size_t n, const uint32_t* idx, const float* val, int nnz, cudaStream_t st) {
  FB_CUDA(cudaMemsetAsync(dst, 0, n * sizeof(float), st));
  if (nnz > 0) {
    int grid = (nnz + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    scatter_kernel<<<grid, 256, 0, st>>>(dst, idx, val, nnz, n);
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

int launch_reduce_partials(const float* lpart, int nparts_l, const float* gpart, int nparts_g, int P, float* out,
                           bool with_grad, cudaStream_t st) {
  const int n = kHeader + (with_grad ? P : 0);
  FinishArgs f{};
  reduce_partials_kernel<<<(n + 31) / 32, 256, 0, st>>>(lpart, nparts_l, gpart, nparts_g, P, out, with_grad ? 1 : 0, f);
  FB_CUDA(cudaGetLastError());
  return 0;
}
// reduce + Adam + counters + loss record in one launch (fbsdej_solver_train_steps)
int launch_reduce_adam(const float* lpart, int nparts_l, const float* gpart, int nparts_g, int P, float* out, float* theta,
                       float* m, float* v, const float* mask, float lr, float b1, float b2, float eps, int* t_dev,
                       uint32_t* iter_dev, float* loss_dst, uint32_t* step_ctr, unsigned int* done_ctr, cudaStream_t st,
                       const XchgArgs* x) {
  const int n = kHeader + P;
  FinishArgs f{theta, m, v, mask, lr, b1, b2, eps, t_dev, iter_dev, loss_dst, step_ctr, done_ctr, XchgArgs{}};
  if (x) f.x = *x;
  reduce_partials_kernel<<<(n + 31) / 32, 256, 0, st>>>(lpart, nparts_l, gpart, nparts_g, P, out, 1, f);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_adam(float* theta, float* m, float* v, const float* grad, const float* mask, int n, float lr, float b1,
                float b2, float eps, int* t_dev, cudaStream_t st) {
  adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(theta, m, v, grad, mask, n, lr, b1, b2, eps, t_dev);
  bump_i32_kernel<<<1, 1, 0, st>>>(t_dev);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_bump_u32(uint32_t* p, cudaStream_t st) {
  bump_u32_kernel<<<1, 1, 0, st>>>(p);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_copy_loss(const float* out, float* dst, uint32_t* ctr, cudaStream_t st) {
  copy_loss_kernel<<<1, 1, 0, st>>>(out, dst, ctr);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_net_forward(const float* theta_net, int nin, int H, int L, int nout, int act, const float* x, int rows,
                       float* y, cudaStream_t st) {
  FB_REQUIRE(H <= 64 && L >= 1, "net_forward: H must be <= 64 and L >= 1");
  const int np = nin * H + H + (L - 1) * (H * H + H) + H * nout + nout;
  int grid = (rows + 127) / 128;
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  net_forward_kernel<<<grid, 128, np * sizeof(float), st>>>(theta_net, nin, H, L, nout, act, x, rows, y);
  FB_CUDA(cudaGetLastError());
  return 0;
}
int launch_transpose(const float* src, float* dst, int N, int B, int d, bool to_ndb, cudaStream_t st) {
  const size_t total = (size_t)N * B * d;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  transpose_kernel<<<(int)g, 256, 0, st>>>(src, dst, N, B, d, to_ndb ? 1 : 0);
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fbsdej

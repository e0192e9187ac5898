// Tile MLP primitives: one CTA = kThreads rows; thread r owns row r.
//
// Replaces the Keras `Dense` stacks of the reference (coupledPricing/Networks.py:6-23,
// coupledMFG/Networks.py:6-46: nin -> H (act) -> H (act) -> nout) and the tf.GradientTape pass through them
// (SolversJumpDiff.py:47-53).
//
// Data layout (shared memory)
//   * activation tiles are column-major  tile[col * RS + row]  (RS = 132): thread r only ever touches row r in
//     the forward / delta passes, so the layers of one MLP evaluation need NO barrier; lanes hit consecutive banks.
//   * the constant-1 trick folds biases into the GEMVs: tile XT has a ones column at col nin, H1/H2 have a ones
//     column at col H (< HP), and the weight blocks carry the bias as one more row / column.
//   * per net, rows HP floats wide:
//       W1  [(nin+1)][HP]  row nin = b1         W2  [(H+1)][HP]  row H = b2        W3T [nout][HP]  col H = b3
//       W2T [H][HP]  (W2T[j][k] = W2[k][j])     W1T [H][HP]  (W1T[j][i] = W1[i][j])      (backward only)
//   * the weight gradient is an outer-product GEMM over the rows of the tile.  Every thread owns one fixed 4x4
//     block (x one row chunk) of one of the three weight matrices and keeps its 16 partial sums IN REGISTERS for
//     the whole kernel (all time steps, all tiles); they are flushed once at the end.
#pragma once
#include "common.cuh"

namespace fbsdej {

constexpr int RS = 132;   // row stride (floats) of the column-major tiles

enum { ACT_TANH = 0, ACT_RELU = 1 };

struct NetRt {     // runtime description of a network
  int nin, H, nout, act;
  int ext_off;     // offset of this net in the external flat parameter vector
};

__host__ __device__ inline int net_ext_params(const NetRt& n) {
  return n.nin * n.H + n.H + n.H * n.H + n.H + n.H * n.nout + n.nout;
}
// smem floats of one net's weight block
__host__ __device__ inline int net_smem_floats(const NetRt& n, int HP, bool bwd) {
  return ((n.nin + 1) + (n.H + 1) + n.nout + (bwd ? 2 * n.H : 0)) * HP;
}

template <int HP>
struct NetView {   // smem views of one net
  const float* W1; const float* W2; const float* W3T; const float* W2T; const float* W1T;
  int nin, H, nout, act;
};

__device__ __forceinline__ float act_fn(float a, int act) { return act == ACT_TANH ? tanhf(a) : fmaxf(a, 0.0f); }
__device__ __forceinline__ float dact_fn(float h, int act) { return act == ACT_TANH ? fmaf(-h, h, 1.0f) : (h > 0.0f ? 1.0f : 0.0f); }

// Cooperative load of one net from the external flat vector (W[in][out] row-major, then b) into smem.
template <int HP>
__device__ __forceinline__ NetView<HP> load_net(float* sw, const float* __restrict__ theta, const NetRt& rt, bool bwd) {
  const int nin = rt.nin, H = rt.H, nout = rt.nout;
  const int o1 = 0, o2 = (nin + 1) * HP, o3 = o2 + (H + 1) * HP, o2t = o3 + nout * HP, o1t = o2t + H * HP;
  const int total = net_smem_floats(rt, HP, bwd);
  for (int i = threadIdx.x; i < total; i += blockDim.x) sw[i] = 0.0f;
  __syncthreads();
  const float* __restrict__ th = theta + rt.ext_off;
  const int n1 = nin * H, n2 = n1 + H, n3 = n2 + H * H, n4 = n3 + H, n5 = n4 + H * nout, n6 = n5 + nout;
  for (int e = threadIdx.x; e < n6; e += blockDim.x) {
    const float v = th[e];
    if (e < n1) {
      const int i = e / H, j = e % H;
      sw[o1 + i * HP + j] = v;
      if (bwd) sw[o1t + j * HP + i] = v;
    } else if (e < n2) {
      sw[o1 + nin * HP + (e - n1)] = v;
    } else if (e < n3) {
      const int k = (e - n2) / H, j = (e - n2) % H;
      sw[o2 + k * HP + j] = v;
      if (bwd) sw[o2t + j * HP + k] = v;
    } else if (e < n4) {
      sw[o2 + H * HP + (e - n3)] = v;
    } else if (e < n5) {
      const int k = (e - n4) / nout, j = (e - n4) % nout;
      sw[o3 + j * HP + k] = v;
    } else {
      sw[o3 + (e - n5) * HP + H] = v;
    }
  }
  __syncthreads();
  NetView<HP> nv;
  nv.W1 = sw + o1; nv.W2 = sw + o2; nv.W3T = sw + o3; nv.W2T = sw + o2t; nv.W1T = sw + o1t;
  nv.nin = nin; nv.H = H; nv.nout = nout; nv.act = rt.act;
  return nv;
}

// a[j] = sum_{k<K} in[k*RS] * W[k*HP + j]   (in = this thread's row of a column-major tile)
template <int HP>
__device__ __forceinline__ void gemv(float (&a)[HP], const float* __restrict__ W, int K, const float* __restrict__ in) {
#pragma unroll
  for (int j = 0; j < HP; ++j) a[j] = 0.0f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float hk = in[k * RS];
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
      const float4 w = ld4(W + k * HP + 4 * j4);
      a[4 * j4] = fmaf(hk, w.x, a[4 * j4]);
      a[4 * j4 + 1] = fmaf(hk, w.y, a[4 * j4 + 1]);
      a[4 * j4 + 2] = fmaf(hk, w.z, a[4 * j4 + 2]);
      a[4 * j4 + 3] = fmaf(hk, w.w, a[4 * j4 + 3]);
    }
  }
}

// Tiles of one CTA.  xt/h1/h2/d1/d2 are HP columns wide, dout/out are NOP columns wide (NOP = 12).
constexpr int NOP = 12;
template <int HP>
struct Tiles {
  float* xt; float* h1; float* h2; float* d1; float* d2; float* dout; float* out;
  __host__ __device__ static constexpr int fwd_floats() { return (2 * HP + NOP) * RS; }           // xt, h1, out
  __host__ __device__ static constexpr int bwd_floats() { return (5 * HP + 2 * NOP) * RS; }
  __device__ void carve(float* base, bool bwd) {
    xt = base; h1 = xt + HP * RS; out = h1 + HP * RS;
    if (bwd) { h2 = out + NOP * RS; d1 = h2 + HP * RS; d2 = d1 + HP * RS; dout = d2 + HP * RS; }
    else { h2 = d1 = d2 = dout = nullptr; }
  }
};

// Constant columns: zero everything, callers then keep cols < nin / < H up to date.  The ones columns are
// (re)written by set_ones() whenever the net evaluated on the tile changes shape.
template <int HP>
__device__ __forceinline__ void zero_tiles(float* base, int nfloats) {
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) base[i] = 0.0f;
  __syncthreads();
}

// Forward of this thread's row.  Inputs: xt cols [0, nin) filled by the caller.  Outputs: t.out cols [0, nout).
// KEEP_H2: also store the second hidden layer (needed by the backward sweep).
template <int HP, bool KEEP_H2>
__device__ __forceinline__ void mlp_fwd(const NetView<HP>& nv, const Tiles<HP>& t, int row) {
  float a[HP];
  float* xt = t.xt + row;
  float* h1 = t.h1 + row;
  xt[nv.nin * RS] = 1.0f;
  gemv<HP>(a, nv.W1, nv.nin + 1, xt);
#pragma unroll
  for (int j = 0; j < HP; ++j)
    if (j < nv.H) h1[j * RS] = act_fn(a[j], nv.act);
  h1[nv.H * RS] = 1.0f;
  gemv<HP>(a, nv.W2, nv.H + 1, h1);
#pragma unroll
  for (int j = 0; j < HP; ++j) a[j] = (j < nv.H) ? act_fn(a[j], nv.act) : ((j == nv.H) ? 1.0f : 0.0f);
  if (KEEP_H2) {
    float* h2 = t.h2 + row;
#pragma unroll
    for (int j = 0; j < HP; ++j) h2[j * RS] = a[j];
  }
  float* out = t.out + row;
  for (int j = 0; j < nv.nout; ++j) {
    const float* __restrict__ w = nv.W3T + j * HP;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < HP / 4; ++k4) {
      const float4 wv = ld4(w + 4 * k4);
      acc0 = fmaf(a[4 * k4], wv.x, acc0); acc1 = fmaf(a[4 * k4 + 1], wv.y, acc1);
      acc0 = fmaf(a[4 * k4 + 2], wv.z, acc0); acc1 = fmaf(a[4 * k4 + 3], wv.w, acc1);
    }
    out[j * RS] = acc0 + acc1;
  }
}

// Delta pass of this thread's row (after mlp_fwd<HP,true> on the same inputs).  Inputs: t.dout cols [0, nout).
// Leaves d2/d1 in the tiles (for the weight gradient) and returns dx[i] = dL/dx_i in a[i], i < nin.
template <int HP>
__device__ __forceinline__ void mlp_delta(const NetView<HP>& nv, const Tiles<HP>& t, int row, float (&a)[HP]) {
  float* d2 = t.d2 + row;
  float* d1 = t.d1 + row;
  const float* h1 = t.h1 + row;
  const float* h2 = t.h2 + row;
  gemv<HP>(a, nv.W3T, nv.nout, t.dout + row);
#pragma unroll
  for (int k = 0; k < HP; ++k) d2[k * RS] = (k < nv.H) ? a[k] * dact_fn(h2[k * RS], nv.act) : 0.0f;
  gemv<HP>(a, nv.W2T, nv.H, d2);
#pragma unroll
  for (int k = 0; k < HP; ++k) d1[k * RS] = (k < nv.H) ? a[k] * dact_fn(h1[k * RS], nv.act) : 0.0f;
  gemv<HP>(a, nv.W1T, nv.H, d1);
}

// ---- weight gradient ----------------------------------------------------------------------------------------
// Thread u of the CTA owns block `blk = u % NB` and row chunk `u / NB` of the block list
//   [ dW1: ceil((nin+1)/4) x HP/4 | dW2: HP/4 x HP/4 | dW3T: ceil(nout/4) x HP/4 ].
template <int HP>
struct WGrad {
  float p[4][4];
  int act_off, del_off;     // float offsets of the two 4-column groups from the tile base (xt)
  int rq0, rq1;             // row-quad range of this thread's chunk
  int type, k0, j0;         // 0: dW1[k][j]  1: dW2[k][j]  2: dW3T[j][k];  -1: idle
  int chunk, S;

  __device__ void init(const NetView<HP>& nv, const Tiles<HP>& t) {
    constexpr int JB = HP / 4;
    const int nb1 = ((nv.nin + 1 + 3) / 4) * JB, nb2 = JB * JB, nb3 = ((nv.nout + 3) / 4) * JB;
    const int NB = nb1 + nb2 + nb3;
    S = kThreads / NB;
    S = S < 1 ? 1 : (S > 4 ? 4 : S);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) p[a][b] = 0.0f;
    const int u = threadIdx.x;
    chunk = u / NB;
    type = -1; k0 = j0 = 0; act_off = del_off = 0; rq0 = rq1 = 0;
    if (chunk >= S) return;
    const int blk = u % NB;
    rq0 = (chunk * (kThreads / 4)) / S;
    rq1 = ((chunk + 1) * (kThreads / 4)) / S;
    if (blk < nb1) {
      type = 0; k0 = 4 * (blk / JB); j0 = 4 * (blk % JB);
      act_off = (int)(t.xt - t.xt) + k0 * RS; del_off = (int)(t.d1 - t.xt) + j0 * RS;
    } else if (blk < nb1 + nb2) {
      const int b2 = blk - nb1;
      type = 1; k0 = 4 * (b2 / JB); j0 = 4 * (b2 % JB);
      act_off = (int)(t.h1 - t.xt) + k0 * RS; del_off = (int)(t.d2 - t.xt) + j0 * RS;
    } else {
      const int b3 = blk - nb1 - nb2;
      type = 2; j0 = 4 * (b3 / JB); k0 = 4 * (b3 % JB);
      act_off = (int)(t.dout - t.xt) + j0 * RS; del_off = (int)(t.h2 - t.xt) + k0 * RS;
    }
  }

  // p[a][b] += sum_rows A[(c+a)][r] * B[(c'+b)][r].  Call between two __syncthreads().
  __device__ __forceinline__ void accumulate(const float* __restrict__ base) {
    if (type < 0) return;
    const float* __restrict__ A = base + act_off;
    const float* __restrict__ Bt = base + del_off;
    for (int rq = rq0; rq < rq1; ++rq) {
      float4 av[4], dv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = ld4(A + a * RS + 4 * rq);
#pragma unroll
      for (int b = 0; b < 4; ++b) dv[b] = ld4(Bt + b * RS + 4 * rq);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          p[a][b] = fmaf(av[a].x, dv[b].x, p[a][b]); p[a][b] = fmaf(av[a].y, dv[b].y, p[a][b]);
          p[a][b] = fmaf(av[a].z, dv[b].z, p[a][b]); p[a][b] = fmaf(av[a].w, dv[b].w, p[a][b]);
        }
    }
  }

  // external flat index (relative to the net) of p[a][b], or -1
  __device__ __forceinline__ int ext_index(const NetView<HP>& nv, int a, int b) const {
    const int nin = nv.nin, H = nv.H, nout = nv.nout;
    if (type == 0) {
      const int k = k0 + a, j = j0 + b;
      if (j >= H || k > nin) return -1;
      return k < nin ? k * H + j : nin * H + j;
    }
    if (type == 1) {
      const int k = k0 + a, j = j0 + b;
      if (j >= H || k > H) return -1;
      const int base = nin * H + H;
      return k < H ? base + k * H + j : base + H * H + j;
    }
    if (type == 2) {
      const int j = j0 + a, k = k0 + b;
      if (j >= nout || k > H) return -1;
      const int base = nin * H + H + H * H + H;
      return k < H ? base + k * nout + j : base + H * nout + j;
    }
    return -1;
  }

  // Add this thread's block into the smem gradient vector sg (external layout), chunk by chunk (deterministic).
  __device__ void flush(const NetView<HP>& nv, float* sg, int ext_off) {
    for (int c = 0; c < 4; ++c) {
      if (type >= 0 && chunk == c) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int e = ext_index(nv, a, b);
            if (e >= 0) sg[ext_off + e] += p[a][b];
          }
      }
      __syncthreads();
    }
  }
};

}  // namespace fbsdej

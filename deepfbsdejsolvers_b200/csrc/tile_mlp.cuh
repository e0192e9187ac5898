// Tile MLP primitives: one CTA = 128 rows; thread r owns row r.
//
// Replaces the Keras `Dense` stacks of the reference (coupledPricing/Networks.py:6-23,
// coupledMFG/Networks.py:6-46: nin -> H (act) -> ... -> H (act) -> nout with L = 1, 2 or 3 equal hidden layers - the
// reference's `--nbLayer`, default 2) and the tf.GradientTape pass through them (SolversJumpDiff.py:47-53).
//
// Data layout (shared memory)
//   * activation tiles use the UMMA canonical K-major (no-swizzle) layout of sm_100: feature f of row r lives at
//     float index ((f/4)*128 + r)*4 + f%4, i.e. float4 chunks [f/4][r].  Thread r reads/writes whole float4 chunks
//     of its own row (conflict-free 128-bit accesses, no barrier between the layers of one evaluation), and the
//     same bytes are directly a tcgen05.mma A operand (8-row x 16-byte core matrices, SBO = 128 B, LBO = 2048 B).
//   * the constant-1 trick folds biases into the GEMVs: tile XT has a ones column at col nin, H1/H2 have a ones
//     column at col H (< HP), and the weight blocks carry the bias as one more row / column.
//   * per net, rows HP floats wide, row counts padded to a multiple of 4 with zero rows:
//       W1  [(nin+1)][HP]  row nin = b1         Wh[l] [(H+1)][HP]  row H = bias (l < L-1)   W3T [nout][HP]  col H = b3
//       WhT[l] [H][HP]  (transposes)            W1T [H][HP]  (W1T[j][i] = W1[i][j])      (backward only)
//   * the weight gradient is an outer-product GEMM over the rows of the tile.  Every thread owns one fixed 4x4
//     block (x one row chunk) of one of the three weight matrices and keeps its 16 partial sums IN REGISTERS for
//     the whole kernel (all time steps, all tiles); they are flushed once at the end.
#pragma once
#include "common.cuh"

namespace fbsdej {

constexpr int TR = 128;   // rows per tile (= kThreads)
constexpr int NOP = 12;   // largest supported nout (width of the out / dout tiles of the jump-scheme kernels)

enum { ACT_TANH = 0, ACT_RELU = 1 };

constexpr int kMaxL = 3;   // hidden layers supported by the fp32 tile MLP (the tcgen05 kernels: 2)

struct NetRt {     // runtime description of a network
  int nin, H, nout, act;
  int ext_off;     // offset of this net in the external flat parameter vector
  int L;           // hidden layers (1 .. kMaxL), all H wide
};

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int net_ext_params(const NetRt& n) {
  return n.nin * n.H + n.H + (n.L - 1) * (n.H * n.H + n.H) + n.H * n.nout + n.nout;
}
// smem floats of one net's weight block
__host__ __device__ inline int net_smem_floats(const NetRt& n, int HP, bool bwd) {
  return (pad4(n.nin + 1) + (n.L - 1) * pad4(n.H + 1) + n.nout + (bwd ? n.L * pad4(n.H) : 0)) * HP;
}

// float index of (feature f, row r) in a tile
__device__ __forceinline__ int tix(int f, int r) { return (((f >> 2) * TR + r) << 2) | (f & 3); }

template <int HP>
struct NetView {   // smem views of one net
  const float* W1; const float* Wh[kMaxL - 1]; const float* W3T; const float* WhT[kMaxL - 1]; const float* W1T;
  int nin, H, nout, act, L;
};

// tanh(x) = 1 - 2 / (2^(2 log2(e) x) + 1): two MUFU ops (ex2, rcp) + three FP32 ops, branch-free, saturates correctly;
// absolute error <= ~2e-7 (the library tanhf costs ~25 instructions with a divergent small/large-argument split).
__device__ __forceinline__ float tanh_fast(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ float act_fn(float a, int act) { return act == ACT_TANH ? tanh_fast(a) : fmaxf(a, 0.0f); }
__device__ __forceinline__ float dact_fn(float h, int act) { return act == ACT_TANH ? fmaf(-h, h, 1.0f) : (h > 0.0f ? 1.0f : 0.0f); }

// Cooperative load of one net from the external flat vector (W[in][out] row-major, then b) into smem.
template <int HP>
__device__ __forceinline__ NetView<HP> load_net(float* sw, const float* __restrict__ theta, const NetRt& rt, bool bwd) {
  const int nin = rt.nin, H = rt.H, nout = rt.nout, L = rt.L;
  const int o1 = 0, oh = pad4(nin + 1) * HP, szh = pad4(H + 1) * HP, o3 = oh + (L - 1) * szh, oht = o3 + nout * HP,
            szt = pad4(H) * HP, o1t = oht + (L - 1) * szt;
  const int total = net_smem_floats(rt, HP, bwd);
  for (int i = threadIdx.x; i < total; i += blockDim.x) sw[i] = 0.0f;
  __syncthreads();
  const float* __restrict__ th = theta + rt.ext_off;
  const int n1 = nin * H, n2 = n1 + H, blk = H * H + H, n4 = n2 + (L - 1) * blk, n5 = n4 + H * nout, n6 = n5 + nout;
  for (int e = threadIdx.x; e < n6; e += blockDim.x) {
    const float v = th[e];
    if (e < n1) {
      const int i = e / H, j = e % H;
      sw[o1 + i * HP + j] = v;
      if (bwd) sw[o1t + j * HP + i] = v;
    } else if (e < n2) {
      sw[o1 + nin * HP + (e - n1)] = v;
    } else if (e < n4) {                                  // hidden-to-hidden layer l: W[k][j] then b[j]
      const int l = (e - n2) / blk, r = (e - n2) % blk;
      if (r < H * H) {
        const int k = r / H, j = r % H;
        sw[oh + l * szh + k * HP + j] = v;
        if (bwd) sw[oht + l * szt + j * HP + k] = v;
      } else {
        sw[oh + l * szh + H * HP + (r - H * H)] = v;
      }
    } else if (e < n5) {
      const int k = (e - n4) / nout, j = (e - n4) % nout;
      sw[o3 + j * HP + k] = v;
    } else {
      sw[o3 + (e - n5) * HP + H] = v;
    }
  }
  __syncthreads();
  NetView<HP> nv;
  nv.W1 = sw + o1; nv.W3T = sw + o3; nv.W1T = sw + o1t;
#pragma unroll
  for (int l = 0; l < kMaxL - 1; ++l) { nv.Wh[l] = sw + oh + l * szh; nv.WhT[l] = sw + oht + l * szt; }
  nv.nin = nin; nv.H = H; nv.nout = nout; nv.act = rt.act; nv.L = L;
  return nv;
}

// one input value against one weight row: a[j] += x * W[j]
template <int HP>
__device__ __forceinline__ void axpy_row(float (&a)[HP], float x, const float* __restrict__ W) {
#pragma unroll
  for (int j4 = 0; j4 < HP / 4; ++j4) {
    const float4 w = ld4(W + 4 * j4);
    a[4 * j4] = fmaf(x, w.x, a[4 * j4]);
    a[4 * j4 + 1] = fmaf(x, w.y, a[4 * j4 + 1]);
    a[4 * j4 + 2] = fmaf(x, w.z, a[4 * j4 + 2]);
    a[4 * j4 + 3] = fmaf(x, w.w, a[4 * j4 + 3]);
  }
}

// a[j] = sum_{k < 4*K4} in[k] * W[k*HP + j]   (in = this thread's row of a tile; W rows beyond the true K are zero)
template <int HP>
__device__ __forceinline__ void gemv(float (&a)[HP], const float* __restrict__ W, int K4, const float* __restrict__ tile, int row) {
#pragma unroll
  for (int j = 0; j < HP; ++j) a[j] = 0.0f;
  const float* __restrict__ in = tile + 4 * row;
#pragma unroll 2
  for (int c = 0; c < K4; ++c) {
    const float4 x = ld4(in + c * (4 * TR));
    const float* __restrict__ w = W + 4 * c * HP;
    axpy_row<HP>(a, x.x, w);
    axpy_row<HP>(a, x.y, w + HP);
    axpy_row<HP>(a, x.z, w + 2 * HP);
    axpy_row<HP>(a, x.w, w + 3 * HP);
  }
}

// store this thread's row of an HP-wide register vector into a tile
template <int HP>
__device__ __forceinline__ void store_row(float* __restrict__ tile, int row, const float (&v)[HP]) {
#pragma unroll
  for (int c = 0; c < HP / 4; ++c) st4(tile + (c * TR + row) * 4, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
}

// Tiles of one CTA.  xt / h[l] / d[l] are HP features wide, dout / out are NW wide (NW = 4 or 12, >= nout).  Order in memory:
// xt | h[0] | out | h[1 .. L-1] | d[0 .. L-1] | dout  (forward-only: xt | h[0] | out | h[1] when L = 3).
template <int HP, int NW>
struct Tiles {
  float* xt; float* h[kMaxL]; float* d[kMaxL]; float* dout; float* out;
  __host__ __device__ static constexpr int fwd_floats(int L = 2) { return ((L > 2 ? 3 : 2) * HP + NW) * TR; }     // xt, h[0] (, h[1]), out
  __host__ __device__ static constexpr int bwd_floats(int L = 2) { return ((1 + 2 * L) * HP + 2 * NW) * TR; }
  __device__ void carve(float* base, bool bwd, int L = 2) {
    xt = base; h[0] = xt + HP * TR; out = h[0] + HP * TR;
    float* p = out + NW * TR;
#pragma unroll
    for (int l = 1; l < kMaxL; ++l) {
      h[l] = nullptr;
      if (l < L && (bwd || l < 2)) { h[l] = p; p += HP * TR; }
    }
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      d[l] = nullptr;
      if (bwd && l < L) { d[l] = p; p += HP * TR; }
    }
    dout = bwd ? p : nullptr;
  }
};

__device__ __forceinline__ void zero_tiles(float* base, int nfloats) {
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) base[i] = 0.0f;
  __syncthreads();
}

// Forward of this thread's row.  Inputs: xt features [0, nin] written by the caller (feature nin = 1.0).
// Outputs: t.out features [0, nout).  KEEP: also store the last hidden layer (needed by the backward sweep).
template <int HP, bool KEEP, class TL>
__device__ __forceinline__ void mlp_fwd(const NetView<HP>& nv, const TL& t, int row) {
  float a[HP];
  gemv<HP>(a, nv.W1, (nv.nin + 4) >> 2, t.xt, row);
#pragma unroll
  for (int j = 0; j < HP; ++j) a[j] = (j < nv.H) ? act_fn(a[j], nv.act) : ((j == nv.H) ? 1.0f : 0.0f);
#pragma unroll
  for (int l = 0; l < kMaxL - 1; ++l) {
    if (l < nv.L - 1) {                                  // a = h[l] -> h[l + 1]
      store_row<HP>(t.h[l], row, a);
      gemv<HP>(a, nv.Wh[l], (nv.H + 4) >> 2, t.h[l], row);
#pragma unroll
      for (int j = 0; j < HP; ++j) a[j] = (j < nv.H) ? act_fn(a[j], nv.act) : ((j == nv.H) ? 1.0f : 0.0f);
    }
  }
  if (KEEP) store_row<HP>(t.h[nv.L - 1], row, a);
  for (int j = 0; j < nv.nout; ++j) {
    const float* __restrict__ w = nv.W3T + j * HP;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < HP / 4; ++k4) {
      const float4 wv = ld4(w + 4 * k4);
      acc0 = fmaf(a[4 * k4], wv.x, acc0); acc1 = fmaf(a[4 * k4 + 1], wv.y, acc1);
      acc0 = fmaf(a[4 * k4 + 2], wv.z, acc0); acc1 = fmaf(a[4 * k4 + 3], wv.w, acc1);
    }
    t.out[tix(j, row)] = acc0 + acc1;
  }
}

// a[j] *= act'(h[j]) for j < H, 0 beyond (h = this row of a hidden tile)
template <int HP>
__device__ __forceinline__ void times_dact(float (&a)[HP], const float* __restrict__ htile, int row, int H, int act) {
#pragma unroll
  for (int c = 0; c < HP / 4; ++c) {
    const float4 h = ld4(htile + (c * TR + row) * 4);
    a[4 * c] = (4 * c < H) ? a[4 * c] * dact_fn(h.x, act) : 0.0f;
    a[4 * c + 1] = (4 * c + 1 < H) ? a[4 * c + 1] * dact_fn(h.y, act) : 0.0f;
    a[4 * c + 2] = (4 * c + 2 < H) ? a[4 * c + 2] * dact_fn(h.z, act) : 0.0f;
    a[4 * c + 3] = (4 * c + 3 < H) ? a[4 * c + 3] * dact_fn(h.w, act) : 0.0f;
  }
}

// Delta pass of this thread's row (after mlp_fwd<HP,true> on the same inputs).  Inputs: t.dout features [0, nout)
// (features up to the next multiple of 4 must be finite).  Leaves d[l] in the tiles (for the weight gradient) and
// returns dx[i] = dL/dx_i in a[i], i < nin.
template <int HP, class TL>
__device__ __forceinline__ void mlp_delta(const NetView<HP>& nv, const TL& t, int row, float (&a)[HP]) {
  // top delta = (W3 dout) .* act'(h[L-1])   -- W3T rows beyond nout belong to the next weight block: mask the inputs instead
  {
#pragma unroll
    for (int j = 0; j < HP; ++j) a[j] = 0.0f;
    for (int j = 0; j < nv.nout; ++j) axpy_row<HP>(a, t.dout[tix(j, row)], nv.W3T + j * HP);
  }
  times_dact<HP>(a, t.h[nv.L - 1], row, nv.H, nv.act);
  store_row<HP>(t.d[nv.L - 1], row, a);
#pragma unroll
  for (int l = kMaxL - 2; l >= 0; --l) {
    if (l < nv.L - 1) {                                  // delta of hidden layer l from the one above
      gemv<HP>(a, nv.WhT[l], (nv.H + 3) >> 2, t.d[l + 1], row);
      times_dact<HP>(a, t.h[l], row, nv.H, nv.act);
      store_row<HP>(t.d[l], row, a);
    }
  }
  gemv<HP>(a, nv.W1T, (nv.H + 3) >> 2, t.d[0], row);
}

// ---- one row shared by the whole CTA ------------------------------------------------------------------------------
// Jump schemes at small batch: the CTA (G == kThreads) works on ONE path, so the network at the path's state has a single
// row.  Thread j owns hidden unit j; the layer vectors live in a small shared-memory block
//   rv: xs[HP] | hs[kMaxL][HP] | ds[kMaxL][HP] | dos[HP] | outs[16]
// L + 1 barriers per forward, L + 1 more per delta pass - instead of every thread walking the same row through the tile MLP.
template <int HP>
__host__ __device__ constexpr int row_floats() { return (2 + 2 * kMaxL) * HP + 16; }
template <int HP> __host__ __device__ constexpr int rv_h(int l) { return (1 + l) * HP; }
template <int HP> __host__ __device__ constexpr int rv_d(int l) { return (1 + kMaxL + l) * HP; }
template <int HP> __host__ __device__ constexpr int rv_do() { return (1 + 2 * kMaxL) * HP; }
template <int HP> __host__ __device__ constexpr int rv_out() { return (2 + 2 * kMaxL) * HP; }

// x: CTA-uniform inputs (x[nin] = 1).  Outputs in rv[rv_out<HP>() + o], o < nout, valid after the call.
template <int HP>
__device__ __forceinline__ void row_fwd(const NetView<HP>& nv, float* __restrict__ rv, const float (&x)[HP]) {
  const int j = threadIdx.x;
  __syncthreads();                                  // readers of the previous row are done
  if (j == 32) {                                    // the inputs, for the weight gradient (a warp with no other work here)
#pragma unroll
    for (int c = 0; c < HP / 4; ++c) st4(rv + 4 * c, make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
  }
  if (j < HP) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < HP; ++i)
      if (i <= nv.nin) acc = fmaf(x[i], nv.W1[i * HP + j], acc);
    rv[rv_h<HP>(0) + j] = (j < nv.H) ? act_fn(acc, nv.act) : ((j == nv.H) ? 1.0f : 0.0f);
  }
  __syncthreads();
  for (int l = 0; l < nv.L - 1; ++l) {
    if (j < HP) {
      float acc = 0.0f;
      const float* __restrict__ hin = rv + rv_h<HP>(l);
      const float* __restrict__ W = nv.Wh[l];
#pragma unroll 8
      for (int k = 0; k <= nv.H; ++k) acc = fmaf(hin[k], W[k * HP + j], acc);
      rv[rv_h<HP>(l + 1) + j] = (j < nv.H) ? act_fn(acc, nv.act) : ((j == nv.H) ? 1.0f : 0.0f);
    }
    __syncthreads();
  }
  if (j < nv.nout) {
    const float* __restrict__ w = nv.W3T + j * HP;
    const float* __restrict__ hin = rv + rv_h<HP>(nv.L - 1);
    float acc = 0.0f;
#pragma unroll 8
    for (int k = 0; k <= nv.H; ++k) acc = fmaf(hin[k], w[k], acc);
    rv[rv_out<HP>() + j] = acc;
  }
  __syncthreads();
}
// dout_j: this thread's entry of dL/dout (thread j < nout; anything elsewhere).  Leaves ds[l] / dos for the weight
// gradient and returns dx[i] = dL/dx_i for i < NDX in every thread.
template <int HP, int NDX>
__device__ __forceinline__ void row_delta(const NetView<HP>& nv, float* __restrict__ rv, float dout_j, float (&dx)[NDX]) {
  const int j = threadIdx.x;
  if (j < HP) rv[rv_do<HP>() + j] = (j < nv.nout) ? dout_j : 0.0f;
  __syncthreads();
  if (j < HP) {
    float acc = 0.0f;
    for (int o = 0; o < nv.nout; ++o) acc = fmaf(rv[rv_do<HP>() + o], nv.W3T[o * HP + j], acc);
    rv[rv_d<HP>(nv.L - 1) + j] = (j < nv.H) ? acc * dact_fn(rv[rv_h<HP>(nv.L - 1) + j], nv.act) : 0.0f;
  }
  __syncthreads();
  for (int l = nv.L - 2; l >= 0; --l) {
    if (j < HP) {
      float acc = 0.0f;
      const float* __restrict__ din = rv + rv_d<HP>(l + 1);
      const float* __restrict__ WT = nv.WhT[l];
#pragma unroll 8
      for (int k = 0; k < nv.H; ++k) acc = fmaf(din[k], WT[k * HP + j], acc);
      rv[rv_d<HP>(l) + j] = (j < nv.H) ? acc * dact_fn(rv[rv_h<HP>(l) + j], nv.act) : 0.0f;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NDX; ++i) dx[i] = 0.0f;
#pragma unroll 8
  for (int k = 0; k < nv.H; ++k) {
    const float d = rv[rv_d<HP>(0) + k];
#pragma unroll
    for (int i = 0; i < NDX; ++i) dx[i] = fmaf(d, nv.W1T[k * HP + i], dx[i]);
  }
}

// ---- weight gradient ----------------------------------------------------------------------------------------
// Thread u of the CTA owns block `blk = u % NB` and row chunk `u / NB` of the block list
//   [ dW1: ceil((nin+1)/4) x HP/4 | dWh[l]: HP/4 x HP/4 for each hidden-to-hidden layer | dW3T: ceil(nout/4) x HP/4 ]
// (at most kThreads blocks: three hidden layers need HP = 24).
// A block is (feature chunk ca of tile A) x (feature chunk cb of tile B): p[a][b] += sum_rows A[4ca+a][r] B[4cb+b][r];
// one float4 per (chunk, row).  The 8 threads of a quarter-warp start at different rows (row rotation), so their
// 128-bit loads fall into different banks although they walk the same row range.
template <int HP>
struct WGrad {
  float p[4][4];
  int a_off, b_off;         // float offsets of the two chunks from the tile base (xt), row 0
  int r0, nr;               // first row and number of rows of this thread's chunk
  int type, k0, j0;         // 0: dW1[k][j]  1 .. L-1: dWh[type-1][k][j]  L: dW3T[j][k];  -1: idle
  int L;
  int chunk, S;

  // u: index of the calling thread among the kThreads threads that share the tile set (threadIdx.x unless a CTA hosts
  // several tile sets, mfg_kernels.cu)
  template <class TL>
  __device__ void init(const NetView<HP>& nv, const TL& t, int u = -1) {
    constexpr int JB = HP / 4;
    L = nv.L;
    const int nb1 = ((nv.nin + 1 + 3) / 4) * JB, nb2 = JB * JB, nb3 = ((nv.nout + 3) / 4) * JB;
    const int NB = nb1 + (L - 1) * nb2 + nb3;
    S = kThreads / NB;
    S = S < 1 ? 1 : (S > 4 ? 4 : S);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) p[a][b] = 0.0f;
    if (u < 0) u = threadIdx.x;
    chunk = u / NB;
    type = -1; k0 = j0 = 0; a_off = b_off = 0; r0 = nr = 0;
    if (chunk >= S) return;
    const int blk = u % NB;
    r0 = (chunk * TR) / S;
    nr = ((chunk + 1) * TR) / S - r0;
    int ta, tb;
    if (blk < nb1) {
      type = 0; k0 = 4 * (blk / JB); j0 = 4 * (blk % JB);
      ta = (int)(t.xt - t.xt) + (k0 >> 2) * (4 * TR); tb = (int)(t.d[0] - t.xt) + (j0 >> 2) * (4 * TR);
    } else if (blk < nb1 + (L - 1) * nb2) {
      const int l = (blk - nb1) / nb2, b2 = (blk - nb1) % nb2;
      type = 1 + l; k0 = 4 * (b2 / JB); j0 = 4 * (b2 % JB);
      ta = (int)(t.h[l] - t.xt) + (k0 >> 2) * (4 * TR); tb = (int)(t.d[l + 1] - t.xt) + (j0 >> 2) * (4 * TR);
    } else {
      const int b3 = blk - nb1 - (L - 1) * nb2;
      type = L; j0 = 4 * (b3 / JB); k0 = 4 * (b3 % JB);
      ta = (int)(t.dout - t.xt) + (j0 >> 2) * (4 * TR); tb = (int)(t.h[L - 1] - t.xt) + (k0 >> 2) * (4 * TR);
    }
    a_off = ta; b_off = tb;
  }

  // Call between two __syncthreads().
  __device__ __forceinline__ void accumulate(const float* __restrict__ base) {
    if (type < 0) return;
    const float* __restrict__ A = base + a_off + 4 * r0;
    const float* __restrict__ Bt = base + b_off + 4 * r0;
    const int rot = threadIdx.x & 7;
    for (int i = 0; i < nr; ++i) {
      int r = i + rot;
      r = r >= nr ? r - nr : r;
      const float4 av = ld4(A + 4 * r), bv = ld4(Bt + 4 * r);
      p[0][0] = fmaf(av.x, bv.x, p[0][0]); p[0][1] = fmaf(av.x, bv.y, p[0][1]);
      p[0][2] = fmaf(av.x, bv.z, p[0][2]); p[0][3] = fmaf(av.x, bv.w, p[0][3]);
      p[1][0] = fmaf(av.y, bv.x, p[1][0]); p[1][1] = fmaf(av.y, bv.y, p[1][1]);
      p[1][2] = fmaf(av.y, bv.z, p[1][2]); p[1][3] = fmaf(av.y, bv.w, p[1][3]);
      p[2][0] = fmaf(av.z, bv.x, p[2][0]); p[2][1] = fmaf(av.z, bv.y, p[2][1]);
      p[2][2] = fmaf(av.z, bv.z, p[2][2]); p[2][3] = fmaf(av.z, bv.w, p[2][3]);
      p[3][0] = fmaf(av.w, bv.x, p[3][0]); p[3][1] = fmaf(av.w, bv.y, p[3][1]);
      p[3][2] = fmaf(av.w, bv.z, p[3][2]); p[3][3] = fmaf(av.w, bv.w, p[3][3]);
    }
  }

  // single-row variant (row_fwd / row_delta above): only the first row chunk's owner of a block accumulates
  __device__ __forceinline__ void accumulate_row(const float* __restrict__ rv) {
    if (type < 0 || chunk != 0) return;
    const float4 av = ld4(rv + (type == 0 ? k0 : type < L ? rv_h<HP>(type - 1) + k0 : rv_do<HP>() + j0));
    const float4 bv = ld4(rv + (type == 0 ? rv_d<HP>(0) + j0 : type < L ? rv_d<HP>(type) + j0 : rv_h<HP>(L - 1) + k0));
    p[0][0] = fmaf(av.x, bv.x, p[0][0]); p[0][1] = fmaf(av.x, bv.y, p[0][1]);
    p[0][2] = fmaf(av.x, bv.z, p[0][2]); p[0][3] = fmaf(av.x, bv.w, p[0][3]);
    p[1][0] = fmaf(av.y, bv.x, p[1][0]); p[1][1] = fmaf(av.y, bv.y, p[1][1]);
    p[1][2] = fmaf(av.y, bv.z, p[1][2]); p[1][3] = fmaf(av.y, bv.w, p[1][3]);
    p[2][0] = fmaf(av.z, bv.x, p[2][0]); p[2][1] = fmaf(av.z, bv.y, p[2][1]);
    p[2][2] = fmaf(av.z, bv.z, p[2][2]); p[2][3] = fmaf(av.z, bv.w, p[2][3]);
    p[3][0] = fmaf(av.w, bv.x, p[3][0]); p[3][1] = fmaf(av.w, bv.y, p[3][1]);
    p[3][2] = fmaf(av.w, bv.z, p[3][2]); p[3][3] = fmaf(av.w, bv.w, p[3][3]);
  }

  // external flat index (relative to the net) of p[a][b], or -1
  __device__ __forceinline__ int ext_index(const NetView<HP>& nv, int a, int b) const {
    const int nin = nv.nin, H = nv.H, nout = nv.nout;
    if (type == 0) {
      const int k = k0 + a, j = j0 + b;
      if (j >= H || k > nin) return -1;
      return k < nin ? k * H + j : nin * H + j;
    }
    if (type < L) {                                   // hidden-to-hidden layer type - 1
      const int k = k0 + a, j = j0 + b;
      if (j >= H || k > H) return -1;
      const int base = nin * H + H + (type - 1) * (H * H + H);
      return k < H ? base + k * H + j : base + H * H + j;
    }
    if (type == L) {
      const int j = j0 + a, k = k0 + b;
      if (j >= nout || k > H) return -1;
      const int base = nin * H + H + (L - 1) * (H * H + H);
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

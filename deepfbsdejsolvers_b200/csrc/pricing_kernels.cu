// Fused pricing-path kernels: one launch walks every path of the batch through all N Euler steps.
//
//   pricing_forward : per-step MLP evaluations (UZ/U net, jump net on the path's own jump, Monte-Carlo
//                     compensator over the shared samples), BSDE update, closed-form coupling A(i,X),
//                     coupled Euler step of X, loss terms; stores the O(1)-per-step state the adjoint needs.
//   pricing_backward: reverse sweep with the hand-derived adjoint, recomputing activations into
//                     shared-memory tiles, weight gradients accumulated in registers (tile_mlp.cuh).
//
// Row mapping: a CTA owns kThreads rows = (kThreads / G) paths x G rows; the G threads of a path share its
// state and split the compensator samples (G = 1: one thread per path).  With one path per CTA (G == kThreads) the
// (U, Z) network has a single row and is evaluated with thread j = hidden unit j (tile_mlp.cuh: row_fwd / row_delta).
//
// Template flag JTC (mma_mode = 1): the jump evaluations - the path's own jump and the compensator samples - run on
// tcgen05 as rows of 128-row tiles (jump_tc.cuh); everything else in these kernels is unchanged.
//
// Reference loss graphs: coupledPricing/SolversJumpDiff.py:22-44 (Global), :86-115/:162-190 (MultiStep1/2),
// :236-269/:315-347 (SumLocal1/2), :391-415 (SumLocalReg), :461-481 (MultiStepReg); SolversPureJump.py same.
#include "pricing.cuh"
#include <type_traits>
#include "cluster.cuh"
#include "jump_tc.cuh"

namespace fbsdej {

// JTC kernels: the tensor-core block of the jump network sits at the start of shared memory (1024-byte aligned)
// (sized for the wider of the two compiled input rows, so that the launch code need not know d)
constexpr int kJtcFwdFloats = (JumpTcFwd<ACT_TANH, 3>::FLOATS + 31) & ~31, kJtcBwdFloats = (JumpTcBwd<ACT_TANH, 3>::FLOATS + 31) & ~31;
template <int D>
__host__ __device__ constexpr int jtc_nxc() { return 2 + 2 * D <= 16 ? 2 : 3; }   // inputs [t, X, jump features, 1] in 16 or 24 features

int launch_reg_tc_backward(int model, int D, const PricingArgs& a, int grid, cudaStream_t st);   // reg_tc_kernels.cu
int launch_reg_tc_forward(int model, int D, const PricingArgs& a, int grid, cudaStream_t st);
size_t reg_tc_backward_smem();
size_t reg_tc_forward_smem(int D);

template <class Model, int HP, bool JUMP, bool JTC>
__global__ void __launch_bounds__(kThreads, JTC ? 4 : 0) pricing_forward(const PricingArgs a) {
  constexpr int D = Model::D;
  static_assert(!JTC || (JUMP && 2 + 2 * D <= 24), "JTC: jump schemes, inputs in at most 24 features");
  constexpr int NXC = jtc_nxc<D>();
  constexpr bool PF = JTC && D == 1;     // compensator samples requested one tile ahead (d = 10: the registers are worth more)
  extern __shared__ __align__(1024) float smem[];
  const bool two = JUMP && !a.one_net;
  float* swA = smem + (JTC ? kJtcFwdFloats : 0);
  float* swB = swA + net_smem_floats(a.netA, HP, false);
  float* red = swB + ((two && !JTC) ? net_smem_floats(a.netB, HP, false) : 0);
  float* rv = red + kRedFloats;                        // single-row block of the (U, Z) network (jump schemes, G == kThreads)
  float* tb = rv + (JUMP ? row_floats<HP>() : 0);
  using TL = Tiles<HP, JUMP ? NOP : 4>;
  TL t;
  NetView<HP> nvA, nvJ;
  const int Lmax = a.netA.L > a.netB.L ? a.netA.L : a.netB.L;
  t.carve(tb, false, Lmax);
  nvA = load_net<HP>(swA, a.theta, a.netA, false);
  nvJ = (two && !JTC) ? load_net<HP>(swB, a.theta, a.netB, false) : nvA;
  zero_tiles(tb, TL::fwd_floats(Lmax));
  JumpTcFwd<ACT_TANH, NXC> jf;
  if constexpr (JTC) jf.init(smem, a.theta, a.netB);

  const int row = threadIdx.x;
  const int G = JUMP ? a.G : 1, ppb = kThreads / G, g = threadIdx.x % G;
  const int C = JUMP ? a.C : 1;                             // CTAs per path (cluster size); C > 1 implies G == kThreads
  const int crank = (JUMP && C > 1) ? (int)cluster_rank() : 0;
  const bool rowmode = JUMP && G == kThreads;               // one path per CTA: the (U, Z) network has a single row
  int cpar = 0;
  const size_t sB = (size_t)a.B;
  const float rdt = a.r * a.dt;
  float lsum = 0.0f;
  const int ntiles = (a.B + ppb - 1) / ppb;
  for (int tile = blockIdx.x / C; tile < ntiles; tile += gridDim.x / C) {
    const int p0 = tile * ppb + threadIdx.x / G;
    const bool valid = p0 < a.B;
    const int p = valid ? p0 : a.B - 1;
    const bool writer = valid && g == 0 && crank == 0;
    float X[D];
#pragma unroll
    for (int k = 0; k < D; ++k) X[k] = a.x0;
    float Y = (a.scheme == SCH_GLOBAL) ? a.theta[a.y0_off] : 0.0f;
    float Cpre = 0.0f;                               // MultiStep: sum_{j<i} toAdd_j
    float yprev = 0.0f, aprev = 0.0f, lloc = 0.0f;   // SumLocal
    for (int i = 0; i < a.N; ++i) {
      const float tf = (a.scheme == SCH_SUMLOCAL && a.stale_time) ? (float)(i == 0 ? 0 : i - 1) : (float)i;
      float dWv[D], Jv[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        dWv[k] = Model::kBrownian ? a.dW[((size_t)i * D + k) * sB + p] : 0.0f;
        Jv[k] = a.J[((size_t)i * D + k) * sB + p];
      }
      if (i + 1 < a.N) {                             // next step's increments -> L2 while this step computes
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (Model::kBrownian) prefetch_l2(a.dW + ((size_t)(i + 1) * D + k) * sB + p);
          prefetch_l2(a.J + ((size_t)(i + 1) * D + k) * sB + p);
        }
      }
      // JTC: the step's sample counts and this thread's first compensator sample are requested before the (U, Z) network runs;
      // inside the tile loop the next sample is always in flight while the current tile is evaluated
      int nnz = 0, n0 = 0;
      float Jn[D];
      auto jmc_load = [&](int m) {
        const int mc = m < a.Mcap ? m : a.Mcap - 1;
#pragma unroll
        for (int k = 0; k < D; ++k) Jn[k] = a.JMC[((size_t)i * D + k) * a.Mcap + mc];
      };
      if constexpr (JTC) {
        nnz = a.jmc_nnz[i]; n0 = a.jmc_n0[i];
        if constexpr (PF) jmc_load(crank * G + g);
      }
      float y_net = 0.0f, zdw = 0.0f;
      if (a.use_netA) {
        float in[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) in[j] = 0.0f;
        in[0] = tf;
#pragma unroll
        for (int k = 0; k < D; ++k) in[1 + k] = X[k];
        in[1 + D] = 1.0f;
        if (rowmode) {
          row_fwd<HP>(nvA, rv, in);
        } else {
          store_row<HP>(t.xt, row, in);
          mlp_fwd<HP, false, TL>(nvA, t, row);
        }
        auto outA = [&](int j) { return rowmode ? rv[rv_out<HP>() + j] : t.out[tix(j, row)]; };
        if (a.has_y) y_net = outA(0);
        if (JUMP && a.has_z) {
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const float z = outA(a.zoff + k);
            zdw = fmaf(z, dWv[k], zdw);
            if (a.trajZ && writer) a.trajZ[((size_t)i * D + k) * sB + p] = z;
          }
        }
      }
      float gam = 0.0f, comp = 0.0f;
      if constexpr (JTC) {
        // rows of the step: compensator samples 0 .. nnz-1, the zero sample (weight n0), the path's own jump
        const int iters = (nnz + 2 + G * C - 1) / (G * C);
        float gc[2] = {0.0f, 0.0f};
        jf.set_time(tf);
        if (G == 1 && !a.one_net && a.jump_sep) {
          // one path per thread: every row of the tile meets the same compensator sample, and the first layer is separable
          // (jump_tc.cuh: preact / eval_sep) - one GEMM for the state part of the step, then one MMA round trip per sample
          float in[HP], pre[24];
          Model::template jump_input<HP>(a, tf, X, Jv, in);
#pragma unroll
          for (int k = 0; k < Model::kJumpSlots; ++k) in[Model::jump_slot0() + k] = 0.0f;
          jf.preact(reinterpret_cast<const float (&)[8 * NXC]>(in), pre);
          const float scale = Model::jump_scale(X);
          float* const ctab = t.h[0];                      // (the (U, Z) network is done with its tiles for this step)
          for (int m0 = 0; m0 <= nnz; m0 += kThreads) {
            __syncthreads();
            jump_sample_parts<Model>(a, i, m0, nnz, ctab);
            __syncthreads();
            const int mend = (nnz + 1 - m0 < kThreads) ? nnz + 1 - m0 : kThreads;
            for (int mm = 0; mm < mend; ++mm) {
              const float y = jf.eval_sep(pre, ctab + mm * 24, scale);
              gc[1] = fmaf(m0 + mm < nnz ? 1.0f : (float)n0, y, gc[1]);
            }
          }
          Model::template jump_input<HP>(a, tf, X, Jv, in);                  // the path's own jump: a full row
          gc[0] = jf.eval(reinterpret_cast<const float (&)[8 * NXC]>(in));
        } else {
        for (int it = 0; it < iters; ++it) {
          const int m = (it * C + crank) * G + g;
          float Jm[D];
          float w = 0.0f;
#pragma unroll
          for (int k = 0; k < D; ++k) Jm[k] = 0.0f;
          if (m < nnz) {
            w = 1.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = PF ? Jn[k] : a.JMC[((size_t)i * D + k) * a.Mcap + m];
          } else if (m == nnz) {
            w = (float)n0;
          } else if (m == nnz + 1) {
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = Jv[k];
          }
          if (PF && it + 1 < iters) jmc_load(m + G * C);
          float in[HP];
          Model::template jump_input<HP>(a, tf, X, Jm, in);
          const float y = jf.eval(reinterpret_cast<const float (&)[8 * NXC]>(in));
          gc[1] = fmaf(w, y, gc[1]);
          if (m == nnz + 1) gc[0] = y;
        }
        }
        group_allsum2(gc[0], gc[1], G, red);
        if (C > 1) cluster_allsum<2>(gc, red, cpar, C);
        gam = gc[0];
        comp = gc[1] / (float)a.M;
      } else if (JUMP) {
        float in[HP];
        Model::template jump_input<HP>(a, tf, X, Jv, in);
        store_row<HP>(t.xt, row, in);
        mlp_fwd<HP, false, TL>(nvJ, t, row);
        gam = t.out[tix(0, row)];
        const int nnz = a.jmc_nnz[i], n0 = a.jmc_n0[i];
        float csum = 0.0f;
        for (int m = crank * G + g; m <= nnz; m += G * C) {
          float Jm[D];
          float w = 1.0f;
          if (m < nnz) {
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = a.JMC[((size_t)i * D + k) * a.Mcap + m];
          } else {
            w = (float)n0;
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = 0.0f;
          }
          if (w != 0.0f) {
            Model::template jump_input<HP>(a, tf, X, Jm, in);
            store_row<HP>(t.xt, row, in);
            mlp_fwd<HP, false, TL>(nvJ, t, row);
            csum = fmaf(w, t.out[tix(0, row)], csum);
          }
        }
        csum = group_allsum(csum, G, red);
        if (C > 1) {
          float cs[1] = {csum};
          cluster_allsum<1>(cs, red, cpar, C);
          csum = cs[0];
        }
        comp = csum / (float)a.M;
      }
      // ---- loss-graph bookkeeping -----------------------------------------------------------
      float Ysel;
      if (a.scheme == SCH_GLOBAL) {
        if (a.trajY && writer) a.trajY[(size_t)i * sB + p] = Y;
        Y = Y - a.dt * (-a.r * Y) + zdw + gam - comp;        // SolversJumpDiff.py:41
        Ysel = Y;                                            // the UPDATED Y feeds oneStepFrom (:43)
      } else {
        if (a.trajY && writer) a.trajY[(size_t)i * sB + p] = y_net;
        const float ai = rdt * y_net + zdw + gam - comp;     // "toAdd" = -dt f(Y) + Z dW + Gam - mean(comp)
        if (a.scheme == SCH_MULTISTEP) {
          if (writer) a.sch1[(size_t)i * sB + p] = y_net - Cpre;   // u_i ; F_i - g = u_i + (sum_all toAdd - g)
          Cpre += ai;
        } else {
          if (i > 0) {
            const float rho = y_net - yprev - aprev;
            lloc = fmaf(rho, rho, lloc);
            if (writer) a.sch1[(size_t)(i - 1) * sB + p] = rho;
          }
          yprev = y_net; aprev = ai;
        }
        Ysel = y_net;
      }
      // ---- coupled Euler step (pricingModels.py:53-54 / :184-185) -----------------------------
      float Ai, dAb;
      if constexpr (JUMP && Model::kBrownian) {
        if (G > 1 && !a.use_atab) Model::eval_A_group(a, i, X, Ai, dAb, G, g, [&](float& u, float& v) { group_allsum2(u, v, G, red); });
        else Model::eval_A(a, i, X, Ai, dAb);
      } else {
        Model::eval_A(a, i, X, Ai, dAb);
      }
      const float diff = Ysel - Ai;
      const float coup = a.aLin * fabsf(diff) * a.dt;
      const float sgn = a.aLin * a.dt * (diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f));
      if (writer) {
        a.aux_s[(size_t)i * sB + p] = sgn;
        a.aux_dA[(size_t)i * sB + p] = dAb;
#pragma unroll
        for (int k = 0; k < D; ++k) a.trajX[((size_t)i * D + k) * sB + p] = X[k];
      }
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const float E = expf(a.drift_dt + a.sig * dWv[k] + Jv[k]);
        if (!JUMP && writer) a.trajE[((size_t)i * D + k) * sB + p] = E;
        X[k] = X[k] * E + coup;
      }
    }
    // ---- terminal condition ---------------------------------------------------------------------
    const float gN = fmaxf(Model::basket(X) - a.K, 0.0f);
    float lpath = 0.0f;
    if (a.scheme == SCH_GLOBAL) {
      const float e = Y - gN;
      lpath = e * e * a.inv_B;
      if (writer) { a.fin[p] = e; if (a.trajY) a.trajY[(size_t)a.N * sB + p] = Y; }
    } else if (a.scheme == SCH_MULTISTEP) {
      // second sweep over the stored u_k: e_k = F_k - g(X_N), loss = mean_k mean_b e_k^2 (SolversJumpDiff.py:115)
      if (writer) {
        const float Dv = Cpre - gN;
        float se = 0.0f, s2 = 0.0f;
        for (int k = 0; k < a.N; ++k) {
          float* const q = a.sch1 + (size_t)k * sB + p;
          const float e = *q + Dv;
          *q = e;
          se += e;
          s2 = fmaf(e, e, s2);
        }
        lpath = s2 * (a.inv_B / (float)a.N);
        a.fin[p] = se;
        if (a.trajY) a.trajY[(size_t)a.N * sB + p] = gN;
      }
    } else {
      const float rho = gN - yprev - aprev;
      lloc = fmaf(rho, rho, lloc);
      lpath = lloc * a.inv_B;
      if (writer) { a.sch1[(size_t)(a.N - 1) * sB + p] = rho; if (a.trajY) a.trajY[(size_t)a.N * sB + p] = gN; }
    }
    if (writer) {
#pragma unroll
      for (int k = 0; k < D; ++k) a.trajX[((size_t)a.N * D + k) * sB + p] = X[k];
      lsum += lpath;
    }
  }
  const float tot = block_sum(lsum, red);
  if (threadIdx.x == 0) {
    a.lpart[blockIdx.x * 4] = tot;
    a.lpart[blockIdx.x * 4 + 1] = 0.0f; a.lpart[blockIdx.x * 4 + 2] = 0.0f; a.lpart[blockIdx.x * 4 + 3] = 0.0f;
  }
  if constexpr (JTC) jf.finish();
  if (JUMP && C > 1) cooperative_groups::this_cluster().sync();   // no CTA leaves while a peer may still read its slots
}

template <class Model, int HP, bool JUMP, bool JTC>
__global__ void __launch_bounds__(kThreads, JTC ? 2 : 0) pricing_backward(const PricingArgs a) {
  constexpr int D = Model::D;
  extern __shared__ __align__(1024) float smem[];
  const bool two = JUMP && !a.one_net;
  float* swA = smem + (JTC ? kJtcBwdFloats : 0);
  float* swB = swA + net_smem_floats(a.netA, HP, true);
  float* red = swB + ((two && !JTC) ? net_smem_floats(a.netB, HP, true) : 0);
  float* rv = red + kRedFloats;
  float* tb = rv + (JUMP ? row_floats<HP>() : 0);
  using TL = Tiles<HP, JUMP ? NOP : 4>;
  TL t;
  NetView<HP> nvA, nvJ;
  WGrad<HP> wgA, wgB;
  const int Lmax = a.netA.L > a.netB.L ? a.netA.L : a.netB.L;
  t.carve(tb, true, Lmax);
  nvA = load_net<HP>(swA, a.theta, a.netA, true);
  nvJ = (two && !JTC) ? load_net<HP>(swB, a.theta, a.netB, true) : nvA;
  zero_tiles(tb, TL::bwd_floats(Lmax));
  wgA.init(nvA, t);
  if (JUMP && !JTC) wgB.init(nvJ, t);
  // JTC: the operand tiles of the tensor-core block live in the h1 .. d2 tiles of the FFMA network (every tile is rewritten
  // in full by whichever phase uses it next; the phases are separated by drain_w / the MMA waits)
  constexpr int NXC = jtc_nxc<D>();
  constexpr bool PF = JTC;               // (two CTAs per SM whatever the register count: always prefetch)
  using JB = JumpTcBwd<ACT_TANH, NXC>;
  static_assert(JB::TILE_FLOATS <= TL::bwd_floats() - (HP + (JUMP ? NOP : 4)) * TR, "operand tiles fit between the input and dout tiles");
  static_assert(!JTC || 1 + D <= JB::NDX, "input gradients of the state come back in one read");
  JB jb;
  if constexpr (JTC) jb.init(smem, t.h[0], a.theta, a.netB);
  float* const ctab = tb + TL::bwd_floats(Lmax);        // JTC: sample parts of the separable first layer (128 x 24)

  const int row = threadIdx.x;
  const int G = JUMP ? a.G : 1, ppb = kThreads / G, g = threadIdx.x % G;
  const int C = JUMP ? a.C : 1;
  const int crank = (JUMP && C > 1) ? (int)cluster_rank() : 0;
  const bool rowmode = JUMP && G == kThreads;
  int cpar = 0;
  const size_t sB = (size_t)a.B;
  const float invB = a.inv_B, invBN = a.inv_B / (float)a.N, rdt = a.r * a.dt;
  float y0g = 0.0f;
  const int ntiles = (a.B + ppb - 1) / ppb;
  for (int tile = blockIdx.x / C; tile < ntiles; tile += gridDim.x / C) {
    const int p0 = tile * ppb + threadIdx.x / G;
    const bool valid = p0 < a.B;
    const int p = valid ? p0 : a.B - 1;
    const float msk = (valid && g == 0 && crank == 0) ? 1.0f : 0.0f;   // rows that count for the main sites
    const float vmsk = valid ? 1.0f : 0.0f;              // rows that count for the compensator
    float X[D], Xbar[D];
#pragma unroll
    for (int k = 0; k < D; ++k) X[k] = a.trajX[((size_t)a.N * D + k) * sB + p];
    float gbar, Ybar = 0.0f, Esum = 0.0f;
    if (a.scheme == SCH_GLOBAL) {
      Ybar = 2.0f * a.fin[p] * invB;
      gbar = -Ybar;
    } else if (a.scheme == SCH_MULTISTEP) {
      Esum = a.fin[p];                                   // sum_k e_k ; d loss / d g = -2/(NB) sum_k e_k
      gbar = -2.0f * Esum * invBN;
    } else {
      gbar = 2.0f * a.sch1[(size_t)(a.N - 1) * sB + p] * invB;
    }
    {
      const float Gb = Model::basket(X);
      const float ind = (Gb - a.K >= 0.0f) ? 1.0f : 0.0f;    // tf.maximum: gradient to the first argument on ties
#pragma unroll
      for (int k = 0; k < D; ++k) Xbar[k] = gbar * ind * ((D == 1) ? 1.0f : Gb / ((float)D * X[k]));
    }
    for (int i = a.N - 1; i >= 0; --i) {
      const float tf = (a.scheme == SCH_SUMLOCAL && a.stale_time) ? (float)(i == 0 ? 0 : i - 1) : (float)i;
      float Jv[D];
      float cY;
      {
        // adjoint of the coupled Euler step X' = X e^{..} + aLin |Ysel - A(i,X)| dt
        const float s_i = a.aux_s[(size_t)i * sB + p], dAb = a.aux_dA[(size_t)i * sB + p];
        float sumXbar = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) sumXbar += Xbar[k];
        cY = sumXbar * s_i;
        float Ev[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {                 // all loads of the step first (independent, one latency)
          X[k] = a.trajX[((size_t)i * D + k) * sB + p];
          if (JUMP) {
            Ev[k] = Model::kBrownian ? a.dW[((size_t)i * D + k) * sB + p] : 0.0f;
            Jv[k] = a.J[((size_t)i * D + k) * sB + p];
          } else {
            Ev[k] = a.trajE[((size_t)i * D + k) * sB + p];
            Jv[k] = 0.0f;
          }
        }
        if (i > 0) {                                  // the next (earlier) step's state -> L2 while this step computes
#pragma unroll
          for (int k = 0; k < D; ++k) {
            prefetch_l2(a.trajX + ((size_t)(i - 1) * D + k) * sB + p);
            if (!JUMP) prefetch_l2(a.trajE + ((size_t)(i - 1) * D + k) * sB + p);
          }
          prefetch_l2(a.aux_s + (size_t)(i - 1) * sB + p);
          prefetch_l2(a.aux_dA + (size_t)(i - 1) * sB + p);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const float E = JUMP ? expf(a.drift_dt + a.sig * Ev[k] + Jv[k]) : Ev[k];
          Xbar[k] = Xbar[k] * E - cY * Model::dA_k(dAb, X[k]);
        }
      }
      // adjoints of the loss graph
      float abar, ybar = 0.0f;
      if (a.scheme == SCH_GLOBAL) {
        Ybar += cY;
        abar = Ybar;
      } else if (a.scheme == SCH_MULTISTEP) {
        const float e = a.sch1[(size_t)i * sB + p];
        abar = 2.0f * Esum * invBN;                       // sum_{k<=i} Fbar_k
        ybar = 2.0f * e * invBN + rdt * abar + cY;
        Esum -= e;
      } else {
        const float rb = 2.0f * a.sch1[(size_t)i * sB + p] * invB;
        const float rbm = (i > 0) ? 2.0f * a.sch1[(size_t)(i - 1) * sB + p] * invB : 0.0f;
        abar = -rb;
        ybar = rbm - rb + rdt * abar + cY;
      }
      float dx[HP];
      float dXacc[JUMP ? D : 1];      // input-gradient sum over this thread's rows (reduced over the path's G threads below)
      if (JUMP) {
#pragma unroll
        for (int k = 0; k < (JUMP ? D : 1); ++k) dXacc[k] = 0.0f;
      }
      int nnz = 0, n0 = 0;
      float Jn[D];
      auto jmc_load = [&](int m) {
        const int mc = m < a.Mcap ? m : a.Mcap - 1;
#pragma unroll
        for (int k = 0; k < D; ++k) Jn[k] = a.JMC[((size_t)i * D + k) * a.Mcap + mc];
      };
      if constexpr (JTC) {
        nnz = a.jmc_nnz[i]; n0 = a.jmc_n0[i];
        if constexpr (PF) jmc_load(crank * G + g);
      }
      if constexpr (JTC) jb.drain_w();   // the last weight-gradient GEMM of the step above still reads the shared tiles
      if (a.use_netA) {
#pragma unroll
        for (int j = 0; j < HP; ++j) dx[j] = 0.0f;
        dx[0] = tf;
#pragma unroll
        for (int k = 0; k < D; ++k) dx[1 + k] = X[k];
        dx[1 + D] = 1.0f;
        if (rowmode) {
          // thread j holds entry j of dL/dout; only the first CTA of the path's cluster counts (the others see zeros)
          const float mrow = (valid && crank == 0) ? 1.0f : 0.0f;
          float dj = (a.has_y && row == 0) ? ybar * mrow : 0.0f;
          if (a.has_z) {
#pragma unroll
            for (int k = 0; k < D; ++k)
              if (row == a.zoff + k) dj = abar * (Model::kBrownian ? a.dW[((size_t)i * D + k) * sB + p] : 0.0f) * mrow;
          }
          row_fwd<HP>(nvA, rv, dx);
          float dxr[1 + D];
          row_delta<HP, 1 + D>(nvA, rv, dj, dxr);
          if (row == 0) {
#pragma unroll
            for (int k = 0; k < D; ++k) dXacc[JUMP ? k : 0] += dxr[1 + k];
          }
          wgA.accumulate_row(rv);
        } else {
          store_row<HP>(t.xt, row, dx);
          mlp_fwd<HP, true, TL>(nvA, t, row);
          for (int j = 0; j < pad4(a.netA.nout); ++j) t.dout[tix(j, row)] = 0.0f;
          if (a.has_y) t.dout[tix(0, row)] = ybar * msk;
          if (JUMP && a.has_z) {
#pragma unroll
            for (int k = 0; k < D; ++k)
              t.dout[tix(a.zoff + k, row)] = abar * (Model::kBrownian ? a.dW[((size_t)i * D + k) * sB + p] : 0.0f) * msk;
          }
          mlp_delta<HP, TL>(nvA, t, row, dx);
#pragma unroll
          for (int k = 0; k < D; ++k) {
            if (JUMP) dXacc[JUMP ? k : 0] += dx[1 + k]; else Xbar[k] += dx[1 + k];
          }
          __syncthreads();
          wgA.accumulate(tb);
          __syncthreads();
        }
      }
      if constexpr (JTC) {
        float (&dXj)[D] = reinterpret_cast<float (&)[D]>(dXacc);
        const float cscale = -abar / (float)a.M * vmsk;
        const int iters = (nnz + 2 + G * C - 1) / (G * C);
        jb.set_time(tf);
        if (G == 1 && !a.one_net && a.jump_sep) {
          // one path per thread, separable first layer (jump_tc.cuh: preact / step_sep / finish_state)
          float pre[24], sumd1[24], dscale = 0.0f;
#pragma unroll
          for (int j = 0; j < 24; ++j) sumd1[j] = 0.0f;
          Model::template jump_input<HP>(a, tf, X, Jv, dx);
#pragma unroll
          for (int k = 0; k < Model::kJumpSlots; ++k) dx[Model::jump_slot0() + k] = 0.0f;
          jb.preact(reinterpret_cast<const float (&)[8 * NXC]>(dx), pre);
          const float scale = Model::jump_scale(X);
          for (int m0 = 0; m0 <= nnz; m0 += kThreads) {
            __syncthreads();
            jump_sample_parts<Model>(a, i, m0, nnz, ctab);
            __syncthreads();
            const int mend = (nnz + 1 - m0 < kThreads) ? nnz + 1 - m0 : kThreads;
            for (int mm = 0; mm < mend; ++mm) {
              const int m = m0 + mm;
              float xj[8 * NXC];
#pragma unroll
              for (int j = 0; j < 8 * NXC; ++j) xj[j] = 0.0f;
#pragma unroll
              for (int k = 0; k < Model::kJumpSlots; ++k)
                xj[Model::jump_slot0() + k] = scale * Model::jump_feature(a, m < nnz ? a.JMC[((size_t)i * D + k) * a.Mcap + m] : 0.0f);
              jb.step_sep(pre, ctab + mm * 24, scale, xj, cscale * (m < nnz ? 1.0f : (float)n0), sumd1, dscale);
            }
          }
          float dn[JB::NDX];
          jb.finish_state(reinterpret_cast<const float (&)[8 * NXC]>(dx), sumd1, dn);
          {
            float Jz[D];
#pragma unroll
            for (int k = 0; k < D; ++k) Jz[k] = 0.0f;
#pragma unroll
            for (int j = 0; j < JB::NDX; ++j) dx[j] = dn[j];
            Model::template jump_input_grad<HP>(a, Jz, dx, dXj);
            if (Model::kScaleIsState) dXj[0] += dscale;
          }
          Model::template jump_input<HP>(a, tf, X, Jv, dx);                    // the path's own jump: a full row
          jb.step(reinterpret_cast<const float (&)[8 * NXC]>(dx), abar * vmsk, dn);
#pragma unroll
          for (int j = 0; j < JB::NDX; ++j) dx[j] = dn[j];
          Model::template jump_input_grad<HP>(a, Jv, dx, dXj);
        } else {
        for (int it = 0; it < iters; ++it) {
          const int m = (it * C + crank) * G + g;
          float Jm[D];
          float dout = 0.0f;
#pragma unroll
          for (int k = 0; k < D; ++k) Jm[k] = 0.0f;
          if (m < nnz) {
            dout = cscale;
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = PF ? Jn[k] : a.JMC[((size_t)i * D + k) * a.Mcap + m];
          } else if (m == nnz) {
            dout = cscale * (float)n0;
          } else if (m == nnz + 1) {                       // the path's own jump
            dout = abar * vmsk;
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = Jv[k];
          }
          if (PF && it + 1 < iters) jmc_load(m + G * C);
          Model::template jump_input<HP>(a, tf, X, Jm, dx);
          float dn[JB::NDX];
          jb.step(reinterpret_cast<const float (&)[8 * NXC]>(dx), dout, dn);
#pragma unroll
          for (int j = 0; j < JB::NDX; ++j) dx[j] = dn[j];
          Model::template jump_input_grad<HP>(a, Jm, dx, dXj);
        }
        }
#pragma unroll
        for (int k = 0; k < D; ++k) dXj[k] = (G == 1) ? dXj[k] : group_allsum(dXj[k], G, red);
        if (C > 1) cluster_allsum<D>(dXj, red, cpar, C);
#pragma unroll
        for (int k = 0; k < D; ++k) Xbar[k] += dXj[k];
      } else if (JUMP) {
        float (&dXj)[D] = reinterpret_cast<float (&)[D]>(dXacc);
        Model::template jump_input<HP>(a, tf, X, Jv, dx);
        store_row<HP>(t.xt, row, dx);
        mlp_fwd<HP, true, TL>(nvJ, t, row);
        for (int j = 0; j < pad4(nvJ.nout); ++j) t.dout[tix(j, row)] = 0.0f;
        t.dout[tix(0, row)] = abar * msk;
        mlp_delta<HP, TL>(nvJ, t, row, dx);
        Model::template jump_input_grad<HP>(a, Jv, dx, dXj);
        __syncthreads();
        if (a.one_net) wgA.accumulate(tb); else wgB.accumulate(tb);
        __syncthreads();
        const int nnz = a.jmc_nnz[i], n0 = a.jmc_n0[i];
        const float cscale = -abar / (float)a.M * vmsk;
        const int iters = (nnz + 1 + G * C - 1) / (G * C);
        for (int it = 0; it < iters; ++it) {
          const int m = (it * C + crank) * G + g;
          float Jm[D];
          float w = 0.0f;
#pragma unroll
          for (int k = 0; k < D; ++k) Jm[k] = 0.0f;
          if (m < nnz) {
            w = 1.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) Jm[k] = a.JMC[((size_t)i * D + k) * a.Mcap + m];
          } else if (m == nnz) {
            w = (float)n0;
          }
          Model::template jump_input<HP>(a, tf, X, Jm, dx);
          store_row<HP>(t.xt, row, dx);
          mlp_fwd<HP, true, TL>(nvJ, t, row);
          t.dout[tix(0, row)] = cscale * w;
          mlp_delta<HP, TL>(nvJ, t, row, dx);
          Model::template jump_input_grad<HP>(a, Jm, dx, dXj);
          __syncthreads();
          if (a.one_net) wgA.accumulate(tb); else wgB.accumulate(tb);
          __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < D; ++k) dXj[k] = (G == 1) ? dXj[k] : group_allsum(dXj[k], G, red);
        if (C > 1) cluster_allsum<D>(dXj, red, cpar, C);
#pragma unroll
        for (int k = 0; k < D; ++k) Xbar[k] += dXj[k];
      }
      if (a.scheme == SCH_GLOBAL) Ybar *= (1.0f + rdt);
    }
    if (a.scheme == SCH_GLOBAL) y0g += Ybar * msk;
  }
  const float y0tot = block_sum(y0g, red);
  // ---- flush: registers -> smem gradient vector (external layout) -> this CTA's row of gpart ----------
  __syncthreads();
  float* sg = tb;
  for (int e = threadIdx.x; e < a.P; e += blockDim.x) sg[e] = 0.0f;
  __syncthreads();
  wgA.flush(nvA, sg, a.netA.ext_off);
  if (two && !JTC) wgB.flush(nvJ, sg, a.netB.ext_off);
  if constexpr (JTC) jb.flush(sg + a.netB.ext_off);
  if (a.scheme == SCH_GLOBAL && threadIdx.x == 0) sg[a.y0_off] = y0tot;
  __syncthreads();
  float* grow = a.gpart + (size_t)blockIdx.x * a.P;
  for (int e = threadIdx.x; e < a.P; e += blockDim.x) grow[e] = sg[e];
  if (JUMP && C > 1) cooperative_groups::this_cluster().sync();
}

// ---- launch glue ---------------------------------------------------------------------------------
template <int HP>
static size_t pricing_smem(const PricingArgs& a, bool backward) {
  if (a.mma_mode == 1 && !a.has_jump) return backward ? reg_tc_backward_smem() : reg_tc_forward_smem(10);
  const bool two = a.has_jump && !a.one_net;
  const bool jtc = a.mma_mode == 1 && a.has_jump;
  const int w = net_smem_floats(a.netA, HP, backward) +
                (jtc ? (backward ? kJtcBwdFloats : kJtcFwdFloats) : two ? net_smem_floats(a.netB, HP, backward) : 0);
  const int L = a.netA.L > a.netB.L ? a.netA.L : a.netB.L;
  const int tl = a.has_jump ? (backward ? Tiles<HP, NOP>::bwd_floats(L) : Tiles<HP, NOP>::fwd_floats(L))
                            : (backward ? Tiles<HP, 4>::bwd_floats(L) : Tiles<HP, 4>::fwd_floats(L));
  return sizeof(float) * (size_t)(w + kRedFloats + (a.has_jump ? row_floats<HP>() : 0) + tl + (jtc && backward ? 24 * kThreads : 0));
}

template <class Model, int HP, bool JUMP, bool JTC = false>
static int launch_one(const PricingArgs& a, int grid, bool backward, cudaStream_t st) {
  const size_t smem = pricing_smem<HP>(a, backward);
  if (smem > 227 * 1024) { set_error("pricing kernels: shared-memory footprint exceeds 227 KB"); return -1; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  const int C = JUMP ? a.C : 1;
  if (C > 1) {                                        // the CTAs of one path form a thread-block cluster
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  if (!backward) {
    auto kern = pricing_forward<Model, HP, JUMP, JTC>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    FB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  } else {
    auto kern = pricing_backward<Model, HP, JUMP, JTC>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    FB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}
// Compiled widths: HP = 24 (H <= 23: every kernel incl. the tcgen05 variants), HP = 32 (H <= 31: fp32 FFMA kernels), HP = 36
// (H <= 35, e.g. the H = 32 variant of SURVEY 8d config 3: the compensator-free solvers only - the 4 x 4 weight-gradient blocks
// of a d = 10 jump network would outnumber the 128 threads of a CTA).
template <class Model, int HP>
static int launch_pair(const PricingArgs& a, int grid, bool backward, cudaStream_t st) {
  if constexpr (HP == 24) {
    if (a.has_jump && a.mma_mode == 1) {   // jump network on tcgen05 (jump_tc.cuh): two-network schemes, d = 1
      return launch_one<Model, HP, true, true>(a, grid, backward, st);
    }
  }
  if (a.mma_mode == 1 && (HP != 24 || a.has_jump)) { set_error("tcgen05 kernels: hidden width <= 22 only"); return -1; }
  if constexpr (HP == 36) {
    if (a.has_jump) { set_error("pricing kernels: hidden width 32..35 is compiled for the compensator-free (Reg) solvers only"); return -1; }
  } else {
    if (a.has_jump) return launch_one<Model, HP, true>(a, grid, backward, st);
  }
  if (a.mma_mode == 1) {   // compensator-free solvers on tcgen05 (reg_tc_kernels.cu)
    const int model = std::is_same<Model, VGModel>::value ? 1 : 0;
    return backward ? launch_reg_tc_backward(model, Model::D, a, grid, st) : launch_reg_tc_forward(model, Model::D, a, grid, st);
  }
  return launch_one<Model, HP, false>(a, grid, backward, st);
}

// A(iStep, X) for n states (component planes X[d][n]); the drop-in MertonJumpModel.A / VGmodel.A.
template <class Model>
__global__ void price_kernel(const PricingArgs a, int iStep, const float* __restrict__ Xin, int n, float* __restrict__ out) {
  constexpr int D = Model::D;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    float X[D];
#pragma unroll
    for (int k = 0; k < D; ++k) X[k] = Xin[(size_t)k * n + p];
    float A, dAb;
    if (iStep >= a.N) A = fmaxf(Model::basket(X) - a.K, 0.0f);   // pricingModels.py:49
    else Model::eval_A(a, iStep, X, A, dAb);
    out[p] = A;
  }
}

int launch_price(int model, int D, const PricingArgs& a, int iStep, const float* X, int n, float* out, cudaStream_t st) {
  int grid = (n + 127) / 128;
  grid = grid < 1 ? 1 : (grid > 148 * 8 ? 148 * 8 : grid);
  if (model == 0 && D == 1) price_kernel<MertonModel<1>><<<grid, 128, 0, st>>>(a, iStep, X, n, out);
  else if (model == 0 && D == 10) price_kernel<MertonModel<10>><<<grid, 128, 0, st>>>(a, iStep, X, n, out);
  else if (model == 1 && D == 1) price_kernel<VGModel><<<grid, 128, 0, st>>>(a, iStep, X, n, out);
  else { set_error("price: unsupported (model, d)"); return -1; }
  FB_CUDA(cudaGetLastError());
  return 0;
}

template <class Model, int HP, bool JUMP, bool JTC = false>
static int occ_one(const PricingArgs& a, bool backward) {
  const size_t smem = pricing_smem<HP>(a, backward);
  int nb = 0;
  cudaError_t e1, e2;
  if (!backward) {
    auto kern = pricing_forward<Model, HP, JUMP, JTC>;
    e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
  } else {
    auto kern = pricing_backward<Model, HP, JUMP, JTC>;
    e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
  }
  if (JTC && e1 == cudaSuccess && e2 == cudaSuccess) {
    // the occupancy calculator answers 1 for kernels that allocate tensor memory; the hardware co-schedules CTAs as long as
    // registers, shared memory and the 512 TMEM columns (128 backward / 96 forward per CTA) allow
    cudaFuncAttributes fa{};
    const cudaError_t e3 = backward ? cudaFuncGetAttributes(&fa, pricing_backward<Model, HP, JUMP, JTC>)
                                    : cudaFuncGetAttributes(&fa, pricing_forward<Model, HP, JUMP, JTC>);
    if (e3 == cudaSuccess && fa.numRegs > 0) {
      const int by_regs = 65536 / (((fa.numRegs + 7) & ~7) * kThreads);
      const int by_smem = (int)((228 * 1024) / (smem + fa.sharedSizeBytes + 1024));
      const int by_tmem = backward ? 512 / (int)JumpTcBwd<ACT_TANH, jtc_nxc<Model::D>()>::NCOLS : 5;
      nb = std::max(1, std::min(std::min(by_regs, by_smem), by_tmem));
    }
  }
  if (e1 != cudaSuccess || e2 != cudaSuccess || nb < 1) {
    if (getenv("FBSDEJ_DEBUG"))
      fprintf(stderr, "[fbsdej] occupancy query failed (%s / %s, nb=%d, smem=%zu)\n", cudaGetErrorString(e1), cudaGetErrorString(e2), nb, smem);
    (void)cudaGetLastError();
    nb = 1;
  }
  if (getenv("FBSDEJ_DEBUG")) fprintf(stderr, "[fbsdej] occupancy(%s, jtc=%d): %d CTAs/SM, smem %zu\n", backward ? "backward" : "forward", (int)JTC, nb, smem);
  return nb;
}
template <class Model>
static int occ_pair24(const PricingArgs& a, bool backward);
template <class Model, int HP>
static int occ_pair(const PricingArgs& a, bool backward) {
  if constexpr (HP != 24) {
    if (a.has_jump) { if constexpr (HP == 36) return 1; else return occ_one<Model, HP, true>(a, backward); }
    return occ_one<Model, HP, false>(a, backward);
  } else {
    return occ_pair24<Model>(a, backward);
  }
}
template <class Model>
static int occ_pair24(const PricingArgs& a, bool backward) {
  constexpr int HP = 24;
  // tcgen05 kernels, by construction: adjoint 4 CTAs per SM (<= 128 registers, 54.9 KB shared memory, 128 TMEM columns),
  // forward 5 (<= 102 registers, 8 KB, 96 columns); the occupancy calculator does not know about TMEM
  if (a.mma_mode == 1 && !a.has_jump) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 4;
  }
  if (a.has_jump && a.mma_mode == 1) {   // the occupancy calculator does not know about TMEM: 128 / 96 of 512 columns per CTA
    return occ_one<Model, HP, true, true>(a, backward);
  }
  if (a.has_jump) return occ_one<Model, HP, true>(a, backward);
  return occ_one<Model, HP, false>(a, backward);
}
// resident CTAs per SM of the kernel that launch_pricing would run
template <int HP>
static int blocks_per_sm_hp(int model, int D, const PricingArgs& a, bool backward) {
  if (model == 0 && D == 1) return occ_pair<MertonModel<1>, HP>(a, backward);
  if (model == 0 && D == 10) return occ_pair<MertonModel<10>, HP>(a, backward);
  if (model == 1 && D == 1) return occ_pair<VGModel, HP>(a, backward);
  return 1;
}
int pricing_blocks_per_sm(int model, int D, int HP, const PricingArgs& a, bool backward) {
  return HP == 24 ? blocks_per_sm_hp<24>(model, D, a, backward) : HP == 32 ? blocks_per_sm_hp<32>(model, D, a, backward)
                                                                            : blocks_per_sm_hp<36>(model, D, a, backward);
}

size_t pricing_smem_bytes(int HP, const PricingArgs& a, bool backward) {
  return HP == 24 ? pricing_smem<24>(a, backward) : HP == 32 ? pricing_smem<32>(a, backward) : pricing_smem<36>(a, backward);
}

// model: 0 = Merton, 1 = VG.  Returns -1 (with message) for shapes that were not compiled in.
template <int HP>
static int launch_hp(int model, int D, const PricingArgs& a, int grid, bool backward, cudaStream_t st) {
  if (model == 0 && D == 1) return launch_pair<MertonModel<1>, HP>(a, grid, backward, st);
  if (model == 0 && D == 10) return launch_pair<MertonModel<10>, HP>(a, grid, backward, st);
  if (model == 1 && D == 1) return launch_pair<VGModel, HP>(a, grid, backward, st);
  return 1;
}
int launch_pricing(int model, int D, int HP, const PricingArgs& a, int grid, bool backward, cudaStream_t st) {
  const int rc = HP == 24 ? launch_hp<24>(model, D, a, grid, backward, st) : HP == 32 ? launch_hp<32>(model, D, a, grid, backward, st)
               : HP == 36 ? launch_hp<36>(model, D, a, grid, backward, st) : 1;
  if (rc <= 0) return rc;
  set_error("pricing kernels: unsupported (model, d, padded width) = (" + std::to_string(model) + ", " +
            std::to_string(D) + ", " + std::to_string(HP) + "); compiled: Merton d in {1,10}, VG d=1, H<=31 (Reg solvers: H<=35)");
  return -1;
}

}  // namespace fbsdej

// Fused kernels of the smart-grid mean-field game (coupled FBSDE with Cox-process jumps).
//
// Reference: coupledMFG/MFGModel.py:35-107 (state update, closed-form controls), coupledMFG/MFGSolvers.py
// loss graphs Global :24-47, MultiStep :187-224, SumLocal :328-364, SumLocalReg :469-505, MultiStepReg :615-651.
// States per path: (hQ, Q, R) exogenous, (hS, S) driven by the networks through the controls.  The
// compensator is analytic (lam*dt), so a step is two small MLP rows plus ~60 flops.  One thread per path.
#include "mfg.cuh"

namespace fbsdej {

struct Controls { float ah, al; };

// calpha_hat / calpha, MFGModel.py:82-89
__device__ __forceinline__ Controls controls(const MFGArgs& a, int i, float hQ, float Q, float R, float hY, float Y) {
  Controls c;
  const float ind = (R <= a.thetaR) ? 1.0f : 0.0f;
  const float ce = a.coeffEqui;
  const float kTheta = a.A + (1.0f - a.pi) * ce * a.p1 + a.K + ce * a.f1 * ind;
  const float mq = a.meanhq[i];
  const float atg = a.stochastic ? a.alphaTarget * mq : a.alphaTarget;
  c.ah = -(1.0f / kTheta) * (a.p0 + a.pi * a.p1 * hQ + ((1.0f - a.pi) * ce * a.p1 + a.K) * hQ + hY +
                             (a.f0 + ce * a.f1 * (hQ - mq - atg)) * ind);
  c.al = -(1.0f / (a.A + a.K)) * (a.K * Q + a.p0 + a.pi * a.p1 * hQ + (1.0f - a.pi) * ce * a.p1 * (hQ + c.ah) + Y +
                                  (a.f0 + ce * a.f1 * (hQ - mq + c.ah - atg)) * ind);
  return c;
}

// Both kernels run TWO roles side by side in a CTA of 2 x 128 threads: threads [0, 128) evaluate / differentiate the
// projected player's network (hat), threads [128, 256) the individual player's, each on its own tile set and for the same
// 128 paths; the network outputs (forward) and the state adjoints (backward) are exchanged through shared memory.  Every
// thread carries the path's scalar state redundantly.  The reference's default batch is ONE tile (B = 128) walking 95 serial
// steps, so the latency of a step - not throughput - is what counts, and the two networks of a step are independent.
constexpr int kMfgThreads = 2 * kThreads;

__device__ __forceinline__ float block_sum2(float v, float* red) {   // sum over the kMfgThreads threads, valid in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

template <int HP>
__global__ void __launch_bounds__(kMfgThreads) mfg_forward(const MFGArgs a) {
  extern __shared__ __align__(16) float smem[];
  using TL = Tiles<HP, 4>;
  float* swA = smem;
  float* swB = swA + net_smem_floats(a.netA, HP, false);
  float* red = swB + net_smem_floats(a.netB, HP, false);
  float* outx = red + 8;                       // network outputs, [parity][role][4][128 rows]
  float* tb = outx + 2 * 2 * 4 * TR;
  const int role = threadIdx.x >> 7, row = threadIdx.x & (TR - 1);
  TL t;
  const int Lmax = a.netA.L > a.netB.L ? a.netA.L : a.netB.L;
  t.carve(tb + role * TL::fwd_floats(Lmax), false, Lmax);
  const NetView<HP> nvA = load_net<HP>(swA, a.theta, a.netA, false);
  const NetView<HP> nvB = load_net<HP>(swB, a.theta, a.netB, false);
  const NetView<HP>& nv = role == 0 ? nvA : nvB;
  zero_tiles(tb, 2 * TL::fwd_floats(Lmax));
  const size_t sB = (size_t)a.B;
  const int c0 = a.has_y ? 1 : 0;
  float lh_sum = 0.0f, li_sum = 0.0f;
  const int ntiles = (a.B + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * TR + row;
    const bool valid = p0 < a.B;
    const bool writer = valid && role == 0;
    const int p = valid ? p0 : a.B - 1;
    float hQ = a.q0, Q = a.q0, R = a.R0, hS = a.S0, S = a.S0;
    float hY = 0.0f, Y = 0.0f;
    if (a.scheme == SCH_GLOBAL) { hY = a.theta[a.y0_off]; Y = a.theta[a.y0_off + 1]; }
    float Ch = 0.0f, Ci = 0.0f;                                                       // MultiStep
    float hyp = 0.0f, ahp = 0.0f, yp = 0.0f, aip = 0.0f, llh = 0.0f, lli = 0.0f;      // SumLocal
    for (int i = 0; i < a.N; ++i) {
      const float tm = (float)i * a.dt;
      const float dW0 = a.dW0[(size_t)i * sB + p], dW = a.dW[(size_t)i * sB + p], dN = a.dN[(size_t)i * sB + p];
      float* const ox = outx + (i & 1) * (2 * 4 * TR);
      t.out = ox + role * (4 * TR);
      if (role == 0) {
        st4(t.xt + 4 * row, make_float4(tm, hQ, hS, R));                             // getProjectedStates
        st4(t.xt + 4 * (TR + row), make_float4(1.0f, 0.0f, 0.0f, 0.0f));
      } else {
        st4(t.xt + 4 * row, make_float4(tm, Q, S, hQ));                              // getAllStates
        st4(t.xt + 4 * (TR + row), make_float4(hS, R, 1.0f, 0.0f));
      }
      mlp_fwd<HP, false, TL>(nv, t, row);
      __syncthreads();                                                               // both networks' outputs of step i
      const float4 oh = ld4(ox + 4 * row), oi = ld4(ox + 4 * TR + 4 * row);
      const float oh0 = oh.x, oh1 = oh.y, oh2 = oh.z;
      const float o0 = oi.x, o1 = oi.y, o2 = oi.z, o3 = oi.w;
      const float lamdt = (a.stochastic ? a.beta * (expf(a.alpha * hQ) - 1.0f) : a.jumpFactor) * a.dt;
      const float dNc = dN - lamdt;
      float a_h = -a.dt * (hS * a.C), a_i = -a.dt * (S * a.C);
      if (a.has_z) {
        const float hz0 = c0 ? oh1 : oh0, hgam = c0 ? oh2 : oh1;
        const float z0 = c0 ? o1 : o0, gam = c0 ? o2 : o1, z = c0 ? o3 : o2;
        a_h = a_h + hz0 * dW0 + hgam * dNc;                    // MFGSolvers.py:40 / :203
        a_i = a_i + z0 * dW0 + gam * dNc + z * dW;             // :41 / :204
      }
      const float hYsel = (a.scheme == SCH_GLOBAL) ? hY : oh0;
      const float Ysel = (a.scheme == SCH_GLOBAL) ? Y : o0;
      if (writer) {
        float* tx = a.traj + ((size_t)i * 5) * sB + p;
        tx[0] = hQ; tx[sB] = Q; tx[2 * sB] = R; tx[3 * sB] = hS; tx[4 * sB] = S;
        if (a.trajY) { a.trajY[((size_t)i * 2) * sB + p] = hYsel; a.trajY[((size_t)i * 2 + 1) * sB + p] = Ysel; }
      }
      if (a.scheme == SCH_GLOBAL) {
        hY += a_h; Y += a_i;
      } else if (a.scheme == SCH_MULTISTEP) {
        if (writer) {
          a.sch[((size_t)i * 2 + 0) * sB + p] = hYsel - Ch;
          a.sch[((size_t)i * 2 + 1) * sB + p] = Ysel - Ci;
        }
        Ch += a_h; Ci += a_i;
      } else {
        if (i > 0) {
          const float rh = hYsel - hyp - ahp, ri = Ysel - yp - aip;
          llh = fmaf(rh, rh, llh); lli = fmaf(ri, ri, lli);
          if (writer) { a.sch[((size_t)(i - 1) * 2 + 0) * sB + p] = rh; a.sch[((size_t)(i - 1) * 2 + 1) * sB + p] = ri; }
        }
        hyp = hYsel; ahp = a_h; yp = Ysel; aip = a_i;
      }
      // oneStepFrom, MFGModel.py:58-71 (controls use the states of step i)
      const Controls c = controls(a, i, hQ, Q, R, hYsel, Ysel);
      hS = hS + c.ah * a.dt;
      S = S + c.al * a.dt;
      R = R + a.dt - (dN > 0.0f ? R : 0.0f);
      const float qn = a.qaver[i + 1];
      hQ = hQ + a.coeffOU * (qn - hQ) * a.dt + a.sig0 * dW0;
      Q = Q + a.coeffOU * (qn - Q) * a.dt + a.sig0 * dW0 + a.sig * dW;
    }
    const float gh = a.h1 + a.h2 * hS, gi = a.h1 + a.h2 * S;
    float lh = 0.0f, li = 0.0f;
    if (a.scheme == SCH_GLOBAL) {
      const float eh = hY - gh, ei = Y - gi;
      lh = eh * eh * a.inv_B; li = ei * ei * a.inv_B;
      if (writer) { a.fin[p] = eh; a.fin[sB + p] = ei; }
    } else if (a.scheme == SCH_MULTISTEP) {
      if (writer) {
        const float Dh = Ch - gh, Di = Ci - gi;
        float seh = 0.0f, sei = 0.0f, s2h = 0.0f, s2i = 0.0f;
        for (int k = 0; k < a.N; ++k) {
          const float eh = a.sch[((size_t)k * 2 + 0) * sB + p] + Dh, ei = a.sch[((size_t)k * 2 + 1) * sB + p] + Di;
          a.sch[((size_t)k * 2 + 0) * sB + p] = eh; a.sch[((size_t)k * 2 + 1) * sB + p] = ei;
          seh += eh; sei += ei; s2h = fmaf(eh, eh, s2h); s2i = fmaf(ei, ei, s2i);
        }
        lh = s2h * (a.inv_B / (float)a.N); li = s2i * (a.inv_B / (float)a.N);
        a.fin[p] = seh; a.fin[sB + p] = sei;
      }
    } else {
      const float rh = gh - hyp - ahp, ri = gi - yp - aip;
      llh = fmaf(rh, rh, llh); lli = fmaf(ri, ri, lli);
      lh = llh * a.inv_B; li = lli * a.inv_B;
      if (writer) { a.sch[((size_t)(a.N - 1) * 2 + 0) * sB + p] = rh; a.sch[((size_t)(a.N - 1) * 2 + 1) * sB + p] = ri; }
    }
    if (writer) {
      float* tx = a.traj + ((size_t)a.N * 5) * sB + p;
      tx[0] = hQ; tx[sB] = Q; tx[2 * sB] = R; tx[3 * sB] = hS; tx[4 * sB] = S;
      if (a.trajY) {
        a.trajY[((size_t)a.N * 2) * sB + p] = (a.scheme == SCH_GLOBAL) ? hY : gh;
        a.trajY[((size_t)a.N * 2 + 1) * sB + p] = (a.scheme == SCH_GLOBAL) ? Y : gi;
      }
      lh_sum += lh; li_sum += li;
    }
  }
  const float th = block_sum2(lh_sum, red);
  const float ti = block_sum2(li_sum, red);
  if (threadIdx.x == 0) {
    a.lpart[blockIdx.x * 4 + 0] = a.w_hat * th + a.w_ind * ti;
    a.lpart[blockIdx.x * 4 + 1] = th;
    a.lpart[blockIdx.x * 4 + 2] = ti;
    a.lpart[blockIdx.x * 4 + 3] = 0.0f;
  }
}

template <int HP>
__global__ void __launch_bounds__(kMfgThreads) mfg_backward(const MFGArgs a) {
  extern __shared__ __align__(16) float smem[];
  using TL = Tiles<HP, 4>;
  float* swA = smem;
  float* swB = swA + net_smem_floats(a.netA, HP, true);
  float* red = swB + net_smem_floats(a.netB, HP, true);
  float* dxx = red + 8;                        // state adjoints out of the two delta passes, [parity][3][128 rows]
  float* tb = dxx + 2 * 3 * TR;
  const int role = threadIdx.x >> 7, row = threadIdx.x & (TR - 1);
  TL t;
  const int Lmax = a.netA.L > a.netB.L ? a.netA.L : a.netB.L;
  float* const tbr = tb + role * TL::bwd_floats(Lmax);
  t.carve(tbr, true, Lmax);
  const NetView<HP> nvA = load_net<HP>(swA, a.theta, a.netA, true);
  const NetView<HP> nvB = load_net<HP>(swB, a.theta, a.netB, true);
  const NetView<HP>& nv = role == 0 ? nvA : nvB;
  zero_tiles(tb, 2 * TL::bwd_floats(Lmax));
  WGrad<HP> wg;
  wg.init(nv, t, row);
  const size_t sB = (size_t)a.B;
  const int c0 = a.has_y ? 1 : 0;
  const float invB = a.inv_B, invBN = a.inv_B / (float)a.N;
  const float wh = a.w_hat, wi = a.w_ind;
  float y0h = 0.0f, y0i = 0.0f;
  const int ntiles = (a.B + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * TR + row;
    const bool valid = p0 < a.B;
    const int p = valid ? p0 : a.B - 1;
    const float msk = valid ? 1.0f : 0.0f;
    float Eh = 0.0f, Ei = 0.0f;
    float ghbar, gibar, hYbar = 0.0f, Ybar = 0.0f;
    if (a.scheme == SCH_GLOBAL) {
      hYbar = 2.0f * a.fin[p] * invB * wh; Ybar = 2.0f * a.fin[sB + p] * invB * wi;
      ghbar = -hYbar; gibar = -Ybar;
    } else if (a.scheme == SCH_MULTISTEP) {
      Eh = a.fin[p]; Ei = a.fin[sB + p];
      ghbar = -2.0f * Eh * invBN * wh;
      gibar = -2.0f * Ei * invBN * wi;
    } else {
      ghbar = 2.0f * a.sch[((size_t)(a.N - 1) * 2 + 0) * sB + p] * invB * wh;
      gibar = 2.0f * a.sch[((size_t)(a.N - 1) * 2 + 1) * sB + p] * invB * wi;
    }
    float hSbar = ghbar * a.h2, Sbar = gibar * a.h2;
    for (int i = a.N - 1; i >= 0; --i) {
      const float tm = (float)i * a.dt;
      const float* tx = a.traj + ((size_t)i * 5) * sB + p;
      const float hQ = tx[0], Q = tx[sB], R = tx[2 * sB], hS = tx[3 * sB], S = tx[4 * sB];
      const float dW0 = a.dW0[(size_t)i * sB + p], dW = a.dW[(size_t)i * sB + p], dN = a.dN[(size_t)i * sB + p];
      const float lamdt = (a.stochastic ? a.beta * (expf(a.alpha * hQ) - 1.0f) : a.jumpFactor) * a.dt;
      const float dNc = dN - lamdt;
      // adjoint of the controlled states: hS' = hS + ah dt, S' = S + al dt
      const float ind = (R <= a.thetaR) ? 1.0f : 0.0f;
      const float ce = a.coeffEqui;
      const float kTheta = a.A + (1.0f - a.pi) * ce * a.p1 + a.K + ce * a.f1 * ind;
      const float albar = Sbar * a.dt;
      const float dal_dah = -(1.0f / (a.A + a.K)) * ((1.0f - a.pi) * ce * a.p1 + ce * a.f1 * ind);
      const float ahbar = hSbar * a.dt + albar * dal_dah;
      const float cYi = albar * (-1.0f / (a.A + a.K));   // adjoint into the Y fed to oneStepFrom
      const float cYh = ahbar * (-1.0f / kTheta);        // adjoint into hY
      float abh, abi, hyb = 0.0f, yb = 0.0f;
      if (a.scheme == SCH_GLOBAL) {
        abh = hYbar; abi = Ybar;                 // hY_{i+1} = hY_i + a_h
        hYbar += cYh; Ybar += cYi;               // OLD hY_i, Y_i feed the controls (MFGSolvers.py:43)
      } else if (a.scheme == SCH_MULTISTEP) {
        const float eh = a.sch[((size_t)i * 2 + 0) * sB + p], ei = a.sch[((size_t)i * 2 + 1) * sB + p];
        abh = 2.0f * Eh * invBN * wh;
        abi = 2.0f * Ei * invBN * wi;
        hyb = 2.0f * eh * invBN * wh + cYh;
        yb = 2.0f * ei * invBN * wi + cYi;
        Eh -= eh; Ei -= ei;
      } else {
        const float rbh = 2.0f * a.sch[((size_t)i * 2 + 0) * sB + p] * invB * wh;
        const float rbi = 2.0f * a.sch[((size_t)i * 2 + 1) * sB + p] * invB * wi;
        float rbhm = 0.0f, rbim = 0.0f;
        if (i > 0) {
          rbhm = 2.0f * a.sch[((size_t)(i - 1) * 2 + 0) * sB + p] * invB * wh;
          rbim = 2.0f * a.sch[((size_t)(i - 1) * 2 + 1) * sB + p] * invB * wi;
        }
        abh = -rbh; abi = -rbi;
        hyb = rbhm - rbh + cYh;
        yb = rbim - rbi + cYi;
      }
      // direct dependence of the increments on the states: a_h = -dt C hS + ..., a = -dt C S + ...
      hSbar += -a.dt * a.C * abh;
      Sbar += -a.dt * a.C * abi;
      float dx[HP];
      float* const dxs = dxx + (i & 1) * (3 * TR);
      if (role == 0) {
        st4(t.xt + 4 * row, make_float4(tm, hQ, hS, R));
        st4(t.xt + 4 * (TR + row), make_float4(1.0f, 0.0f, 0.0f, 0.0f));
        mlp_fwd<HP, true, TL>(nv, t, row);
        float dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (a.has_y) dd[0] = hyb * msk;
        if (a.has_z) {
          const float z0b = abh * dW0 * msk, gb = abh * dNc * msk;
          if (c0) { dd[1] = z0b; dd[2] = gb; } else { dd[0] = z0b; dd[1] = gb; }
        }
        st4(t.dout + 4 * row, make_float4(dd[0], dd[1], dd[2], dd[3]));
        mlp_delta<HP, TL>(nv, t, row, dx);
        dxs[row] = dx[2];                            // d / d hS through the projected player's network
      } else {
        st4(t.xt + 4 * row, make_float4(tm, Q, S, hQ));
        st4(t.xt + 4 * (TR + row), make_float4(hS, R, 1.0f, 0.0f));
        mlp_fwd<HP, true, TL>(nv, t, row);
        float dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (a.has_y) dd[0] = yb * msk;
        if (a.has_z) {
          const float z0b = abi * dW0 * msk, gb = abi * dNc * msk, zb = abi * dW * msk;
          if (c0) { dd[1] = z0b; dd[2] = gb; dd[3] = zb; } else { dd[0] = z0b; dd[1] = gb; dd[2] = zb; }
        }
        st4(t.dout + 4 * row, make_float4(dd[0], dd[1], dd[2], dd[3]));
        mlp_delta<HP, TL>(nv, t, row, dx);
        dxs[TR + row] = dx[2];                       // d / d S
        dxs[2 * TR + row] = dx[4];                   // d / d hS through the individual player's network
      }
      __syncthreads();                               // tiles complete (weight gradient) + adjoints exchanged
      wg.accumulate(tbr);
      hSbar += dxs[row] + dxs[2 * TR + row];
      Sbar += dxs[TR + row];
      __syncthreads();                               // before the tiles are overwritten
    }
    if (a.scheme == SCH_GLOBAL && role == 0) { y0h += hYbar * msk; y0i += Ybar * msk; }
  }
  const float t0 = block_sum2(y0h, red);
  const float t1 = block_sum2(y0i, red);
  __syncthreads();
  float* sg = tb;
  for (int e = threadIdx.x; e < a.P; e += blockDim.x) sg[e] = 0.0f;
  __syncthreads();
  wg.flush(nv, sg, role == 0 ? a.netA.ext_off : a.netB.ext_off);
  if (a.scheme == SCH_GLOBAL && threadIdx.x == 0) { sg[a.y0_off] = t0; sg[a.y0_off + 1] = t1; }
  __syncthreads();
  float* grow = a.gpart + (size_t)blockIdx.x * a.P;
  for (int e = threadIdx.x; e < a.P; e += blockDim.x) grow[e] = sg[e];
}

template <int HP>
static size_t mfg_smem(const MFGArgs& a, bool backward) {
  const int w = net_smem_floats(a.netA, HP, backward) + net_smem_floats(a.netB, HP, backward);
  const int L = a.netA.L > a.netB.L ? a.netA.L : a.netB.L;
  const int tl = backward ? 2 * Tiles<HP, 4>::bwd_floats(L) + 2 * 3 * TR : 2 * Tiles<HP, 4>::fwd_floats(L) + 2 * 2 * 4 * TR;
  return sizeof(float) * (size_t)(w + 8 + tl);
}
size_t mfg_smem_bytes(int HP, const MFGArgs& a, bool backward) {
  return HP == 24 ? mfg_smem<24>(a, backward) : mfg_smem<32>(a, backward);
}

int mfg_blocks_per_sm(int HP, const MFGArgs& a, bool backward) {
  if (HP != 24) return 1;                   // (HP = 32: 199 KB of shared memory, one CTA per SM)
  const size_t smem = mfg_smem<24>(a, backward);
  int nb = 0;
  if (!backward) {
    auto kern = mfg_forward<24>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kMfgThreads, smem) != cudaSuccess) return 1;
  } else {
    auto kern = mfg_backward<24>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kMfgThreads, smem) != cudaSuccess) return 1;
  }
  return nb < 1 ? 1 : nb;
}

template <int HP>
static int launch_mfg_hp(const MFGArgs& a, int grid, bool backward, cudaStream_t st) {
  const size_t smem = mfg_smem<HP>(a, backward);
  if (smem > 227 * 1024) { set_error("mfg kernels: shared-memory footprint exceeds 227 KB"); return -1; }
  if (!backward) {
    auto kern = mfg_forward<HP>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kMfgThreads, smem, st>>>(a);
  } else {
    auto kern = mfg_backward<HP>;
    FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kMfgThreads, smem, st>>>(a);
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}
// compiled widths: HP = 24 (H <= 23), HP = 32 (H <= 31)
int launch_mfg(int HP, const MFGArgs& a, int grid, bool backward, cudaStream_t st) {
  if (HP == 24) return launch_mfg_hp<24>(a, grid, backward, st);
  if (HP == 32) return launch_mfg_hp<32>(a, grid, backward, st);
  set_error("mfg kernels: padded hidden width " + std::to_string(HP) + " not compiled (H <= 31)");
  return -1;
}

}  // namespace fbsdej

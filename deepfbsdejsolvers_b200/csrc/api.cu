// C-ABI of the fbsdej library (include/fbsdej.h): contexts, solver objects, host-side table construction
// (Merton series coefficients, VG Lewis-FFT spline, Poisson inversion table, MFG mean curve), launch glue
// and the CUDA-graph training loop.  No torch types, no C++ exceptions across the boundary.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <complex>
#include <memory>
#include <cstring>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler injects its library

// NVTX range around the launches of one phase of the training path (visible in Nsight Systems / ncu --nvtx); enabled with
// FBSDEJ_NVTX=1 so that the default path pays nothing (tracing).
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) {
    static const bool enabled = getenv("FBSDEJ_NVTX") != nullptr;
    on = enabled;
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

#include "../../include/fbsdej.h"
#include "mfg.cuh"
#include "sim.cuh"
#include "util.cuh"

namespace fbsdej {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace fbsdej

namespace fbsdej {
}
using namespace fbsdej;

struct fbsdej_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  int sms = 148;
};

namespace {

template <class T>
int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) return 0;
  FB_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return 0;
}
template <class T>
int dev_upload(T** p, const std::vector<T>& v, cudaStream_t st) {
  if (dev_alloc(p, v.size())) return -2;
  if (!v.empty()) {
    FB_CUDA(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}
template <class T>
void dev_free(T*& p) {
  if (p) cudaFree((void*)p);
  p = nullptr;
}

int padded_width(int H) { return H <= 23 ? 24 : (H <= 31 ? 32 : (H <= 35 ? 36 : -1)); }

// ---- VG: Lewis-FFT integral table + local not-a-knot cubic spline (pricingModels.py:156-179) -------------
void fft_inplace(std::vector<std::complex<double>>& x, bool inverse) {
  const size_t n = x.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(x[i], x[j]);
  }
  std::vector<std::complex<double>> tw(n / 2);
  for (size_t k = 0; k < n / 2; ++k) {
    const double a = 2.0 * M_PI * (double)k / (double)n * (inverse ? 1.0 : -1.0);
    tw[k] = std::complex<double>(std::cos(a), std::sin(a));
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t stride = n / len;
    for (size_t i = 0; i < n; i += len) {
      for (size_t k = 0; k < len / 2; ++k) {
        const std::complex<double> u = x[i + k], v = x[i + k + len / 2] * tw[k * stride];
        x[i + k] = u + v;
        x[i + k + len / 2] = u - v;
      }
    }
  }
}

// coef: [N][nint][4] cubic coefficients of I_i(k) on [k0 + j h, k0 + (j+1) h]
void build_vg_table(const fbsdej_vg_params& v, int half, std::vector<double>& coef, double& k0, double& h) {
  const int fftN = 1 << 15;
  const double Bq = 500.0, du = Bq / fftN, lm = 2.0 * M_PI / Bq, b = fftN * lm / 2.0;
  const double corr = -std::log(1.0 - v.theta * v.kappa - v.kappa / 2.0 * v.sigJ * v.sigJ) / v.kappa;
  const int margin = 24, c = fftN / 2;            // ku[c] = 0
  const int lo = c - half - margin, n = 2 * (half + margin) + 1;
  h = lm;
  k0 = -b + lm * (double)(c - half);
  const int nint = 2 * half;
  coef.assign((size_t)v.N * nint * 4, 0.0);
  std::vector<std::complex<double>> x(fftN);
  std::vector<double> y(n), Mv(n), rhs(n), cp(n), dp(n);
  const std::complex<double> I(0.0, 1.0);
  for (int i = 0; i < v.N; ++i) {
    const double tau = v.T - i * (v.T / v.N);
    for (int j = 0; j < fftN; ++j) {
      const double u = j * du;
      double w = (j % 2 == 0) ? 2.0 : 4.0;          // 3 + (-1)^(j+1)
      if (j == 0 || j == fftN - 1) w = 1.0;
      const std::complex<double> uc(u, -0.5);
      const std::complex<double> phi =
          std::exp(tau * (I * (v.r - corr) * uc - std::log(1.0 - I * v.theta * v.kappa * uc + 0.5 * v.kappa * v.sigJ * v.sigJ * uc * uc) / v.kappa));
      x[j] = std::exp(-I * (b * j * du)) * phi / (u * u + 0.25) * w * du / 3.0;
    }
    fft_inplace(x, true);                          // = ifft(x) * fftN
    for (int j = 0; j < n; ++j) y[j] = x[lo + j].real();
    // not-a-knot cubic spline on the window (uniform h): second derivatives Mv
    for (int j = 1; j < n - 1; ++j) rhs[j] = 6.0 * (y[j + 1] - 2.0 * y[j] + y[j - 1]) / (h * h);
    Mv[1] = rhs[1] / 6.0;
    Mv[n - 2] = rhs[n - 2] / 6.0;
    // tridiagonal M_{j-1} + 4 M_j + M_{j+1} = rhs_j for j = 2 .. n-3 with known M_1, M_{n-2}
    const int a0 = 2, a1 = n - 3;
    for (int j = a0; j <= a1; ++j) {
      double r = rhs[j];
      if (j == a0) r -= Mv[1];
      if (j == a1) r -= Mv[n - 2];
      const double lower = (j == a0) ? 0.0 : 1.0;
      const double denom = 4.0 - lower * (j == a0 ? 0.0 : cp[j - 1]);
      cp[j] = 1.0 / denom;
      dp[j] = (r - lower * (j == a0 ? 0.0 : dp[j - 1])) / denom;
    }
    for (int j = a1; j >= a0; --j) Mv[j] = dp[j] - (j == a1 ? 0.0 : cp[j] * Mv[j + 1]);
    Mv[0] = 2.0 * Mv[1] - Mv[2];
    Mv[n - 1] = 2.0 * Mv[n - 2] - Mv[n - 3];
    for (int q = 0; q < nint; ++q) {
      const int j = margin + q;
      double* cc = &coef[((size_t)i * nint + q) * 4];
      cc[0] = y[j];
      cc[1] = (y[j + 1] - y[j]) / h - h * (2.0 * Mv[j] + Mv[j + 1]) / 6.0;
      cc[2] = Mv[j] / 2.0;
      cc[3] = (Mv[j + 1] - Mv[j]) / (6.0 * h);
    }
  }
}

}  // namespace

struct fbsdej_solver {
  fbsdej_ctx* ctx = nullptr;
  fbsdej_solver_desc desc{};
  fbsdej_merton_params mer{};
  fbsdej_vg_params vg{};
  fbsdej_mfg_params mfg{};
  int model = 0, D = 1, N = 0, HP = 24, P = 0, y0_off = 0, M = 0;
  double q0 = 0.0;                 // QAver[0]
  long long launches_per_step = 0;
  NetRt netA{}, netB{};
  int sch = 0, one_net = 0, has_jump = 0, use_netA = 1, has_y = 0, zoff = 0, has_z = 0, feat_mode = 0;
  // tables
  float4* tabA = nullptr; float* tabK = nullptr; int2* tab_range = nullptr; float* qdisc = nullptr; int limit = 0;
  float4* atab = nullptr; float4* atab_meta = nullptr; int* atab_off = nullptr; int use_atab = 0;
  float4* vg_coef = nullptr; float* vg_scale = nullptr; int vg_nint = 0; float vg_k0 = 0, vg_h = 1;
  uint32_t* pois_thr = nullptr; int npois = 0;
  float* qaver = nullptr; float* meanhq = nullptr;
  float drift_dt = 0.0f;
  // noise (owned buffers + current pointers)
  int capB = 0, noiseB = 0;
  float *nA = nullptr, *nB = nullptr, *nC = nullptr;
  const float *curA = nullptr, *curB = nullptr, *curC = nullptr;
  float* jmc_raw = nullptr; float* jmc = nullptr; int* jmc_nnz = nullptr; int* jmc_n0 = nullptr;
  // work buffers
  float *trajE = nullptr, *trajX = nullptr, *aux_s = nullptr, *aux_dA = nullptr, *sch1 = nullptr, *fin = nullptr;
  float *rec = nullptr, *recN = nullptr;   // tcgen05 path: tile-major records (pricing.cuh: RecLayout)
  float* wimg = nullptr;                   // tcgen05 compensator-free kernels: weight operand images (reg_tc_kernels.cu)
  float *lpart = nullptr, *gpart = nullptr; int cap_grid = 0;
  float* out_dev = nullptr;       // [4 + P] scratch for train_steps
  uint32_t* step_ctr = nullptr;   // device: [0] step index inside train_steps, [1] finished-block counter of the fused finish
  struct Finish { float* theta; float* m; float* v; const float* mask; float lr, b1, b2, eps; int* t_dev; uint32_t* iter_dev;
                  float* loss_dst; };
  const Finish* finish = nullptr; // set by train_steps: run_pass fuses reduce + Adam + counters into one launch
  // set by step_pass: the tcgen05 forward draws the Merton increments itself (sim_device.cuh), nothing is materialised
  struct Rng { uint64_t seed; uint32_t iteration; const uint32_t* iter_ptr; uint32_t path_offset; };
  const Rng* rng = nullptr;
  // cached training graph
  cudaGraphExec_t graph = nullptr;
  struct Key { const void *theta, *m, *v, *mask, *t, *it, *loss; uint64_t seed; int B; float lr, b1, b2, eps;
               int B_global; uint32_t path_offset; int dp; } key;
  // data-parallel exchange inside the finishing kernel (fbsdej_solver_dp_*, util.cuh: XchgArgs)
  struct Dp {
    int rank = 0, world = 1;
    unsigned char* buf = nullptr;              // this rank's exchange buffer
    std::vector<void*> peers;                  // [world] buffers of every rank as seen from this process
    std::vector<char> opened;                  // peers[r] came from cudaIpcOpenMemHandle
    float** d_data = nullptr; uint32_t** d_flags = nullptr;   // device pointer tables
    bool connected = false;
  } dp;
  bool dp_step = false;                        // set by train_steps_dp around the step
};

namespace {

int free_path_buffers(fbsdej_solver* s) {
  dev_free(s->nA); dev_free(s->nB); dev_free(s->nC);
  dev_free(s->trajE); dev_free(s->trajX); dev_free(s->aux_s); dev_free(s->aux_dA); dev_free(s->sch1); dev_free(s->fin);
  dev_free(s->rec); dev_free(s->recN);
  s->capB = 0;
  return 0;
}

int ensure_capacity(fbsdej_solver* s, int B) {
  if (B <= s->capB) return 0;
  if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
  FB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  free_path_buffers(s);
  const size_t N = s->N, D = s->D, b = B;
  if (s->model == FBSDEJ_MODEL_MFG) {
    if (dev_alloc(&s->nA, N * b) || dev_alloc(&s->nB, N * b) || dev_alloc(&s->nC, N * b)) return -2;
    if (dev_alloc(&s->trajX, (N + 1) * 5 * b) || dev_alloc(&s->sch1, N * 2 * b) || dev_alloc(&s->fin, 2 * b)) return -2;
  } else {
    if (s->model == FBSDEJ_MODEL_MERTON && dev_alloc(&s->nA, N * D * b)) return -2;
    if (dev_alloc(&s->nB, N * D * b)) return -2;
    if (s->desc.mma_mode == 1 && !s->has_jump) {
      const size_t nt = (size_t)make_tile_map(B, 4 * s->ctx->sms).ntiles;
      if (dev_alloc(&s->rec, nt * N * (2 * D + 3) * kThreads) || dev_alloc(&s->recN, nt * (D + 1) * kThreads)) return -2;
    } else {
      if (!s->has_jump && dev_alloc(&s->trajE, N * D * b)) return -2;
      if (dev_alloc(&s->trajX, (N + 1) * D * b) || dev_alloc(&s->aux_s, N * b) || dev_alloc(&s->aux_dA, N * b) ||
          dev_alloc(&s->sch1, N * b) || dev_alloc(&s->fin, b))
        return -2;
    }
  }
  s->capB = B;
  return 0;
}

int ensure_grid(fbsdej_solver* s, int grid) {
  if (grid <= s->cap_grid) return 0;
  if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
  FB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  dev_free(s->lpart); dev_free(s->gpart);
  if (dev_alloc(&s->lpart, (size_t)grid * 4) || dev_alloc(&s->gpart, (size_t)grid * s->P)) return -2;
  s->cap_grid = grid;
  return 0;
}

int pick_G(const fbsdej_solver* s, int B) {
  if (!s->has_jump) return 1;
  const long long target = (long long)s->ctx->sms * kThreads * 2;
  int G = 1;
  while (G < kThreads && (long long)B * G < target) G <<= 1;
  if (G == 64) G = 128;
  return G;
}

// CTAs per path (thread-block cluster) for the jump schemes at small batch: as many as fill the GPU (at most 16),
// but no more than the expected evaluated compensator rows per step give one row per thread.
int pick_C(const fbsdej_solver* s, int B, int G) {
  if (!s->has_jump || G != kThreads || getenv("FBSDEJ_NO_CLUSTER")) return 1;
  double rows = s->M;
  if (s->model == FBSDEJ_MODEL_MERTON)   // zero-jump samples are deduplicated: expected non-zero ones + 1
    rows = s->M * (1.0 - std::exp(-s->mer.lam * (s->mer.T / s->mer.N) * s->D)) + 1.0;
  // (the tcgen05 jump kernels keep two CTAs per SM resident in the adjoint sweep, four in the forward one)
  const int per_sm = s->desc.mma_mode == 1 ? 2 : 1;
  // 16 CTAs per cluster is sm_100's non-portable maximum (the launch sets cudaFuncAttributeNonPortableClusterSizeAllowed);
  // FBSDEJ_MAX_CLUSTER=8 restores the portable limit
  const char* mc = getenv("FBSDEJ_MAX_CLUSTER");
  const int maxC = mc ? std::max(1, std::min(16, atoi(mc))) : 16;
  int C = std::min(maxC, s->ctx->sms * per_sm / std::max(B, 1));
  C = std::min(C, (int)std::ceil(rows / kThreads));
  return std::max(C, 1);
}

void fill_pricing_args(const fbsdej_solver* s, const float* theta, int B, int B_global, PricingArgs& a) {
  std::memset(&a, 0, sizeof(a));
  a.B = B; a.N = s->N; a.G = pick_G(s, B); a.M = s->M > 0 ? s->M : 1;
  a.C = pick_C(s, B, a.G);
  a.scheme = s->sch; a.one_net = s->one_net; a.has_jump = s->has_jump; a.use_netA = s->use_netA;
  a.has_y = s->has_y; a.zoff = s->zoff; a.has_z = s->has_z; a.feat_mode = s->feat_mode;
  a.stale_time = s->desc.stale_time;
  a.mma_mode = s->desc.mma_mode;
  a.jump_sep = getenv("FBSDEJ_NO_JUMP_SEP") ? 0 : 1;
  a.inv_B = 1.0f / (float)B_global;
  if (s->model == FBSDEJ_MODEL_MERTON) {
    a.dt = (float)(s->mer.T / s->mer.N); a.r = (float)s->mer.r; a.K = (float)s->mer.K; a.x0 = (float)s->mer.x0;
    a.aLin = (float)s->mer.aLin; a.sig = (float)s->mer.sig;
  } else {
    a.dt = (float)(s->vg.T / s->vg.N); a.r = (float)s->vg.r; a.K = (float)s->vg.K; a.x0 = (float)s->vg.x0;
    a.aLin = (float)s->vg.aLin; a.sig = 0.0f;
  }
  a.drift_dt = s->drift_dt;
  a.netA = s->netA; a.netB = s->netB; a.y0_off = s->y0_off; a.P = s->P;
  a.theta = theta;
  a.dW = s->curA; a.J = s->curB;
  a.JMC = s->jmc; a.jmc_nnz = s->jmc_nnz; a.jmc_n0 = s->jmc_n0; a.Mcap = s->M > 0 ? s->M : 1;
  a.tabA = s->tabA; a.tabK = s->tabK; a.tab_range = s->tab_range; a.qdisc = s->qdisc; a.limit = s->limit;
  a.atab = s->atab; a.atab_meta = s->atab_meta; a.atab_off = s->atab_off; a.use_atab = s->use_atab;
  a.vg_coef = s->vg_coef; a.vg_scale = s->vg_scale; a.vg_nint = s->vg_nint;
  a.vg_k0 = s->vg_k0; a.vg_h = s->vg_h; a.vg_inv_h = 1.0f / s->vg_h;
  a.trajE = s->trajE; a.trajX = s->trajX; a.aux_s = s->aux_s; a.aux_dA = s->aux_dA; a.sch1 = s->sch1; a.fin = s->fin;
  a.rec = s->rec; a.recN = s->recN;
  a.lpart = s->lpart; a.gpart = s->gpart;
}

void fill_mfg_args(const fbsdej_solver* s, const float* theta, int B, int B_global, MFGArgs& a) {
  std::memset(&a, 0, sizeof(a));
  const fbsdej_mfg_params& m = s->mfg;
  a.B = B; a.N = s->N; a.scheme = s->sch; a.has_y = s->has_y; a.has_z = s->has_z;
  a.stochastic = m.stochastic_jumps;
  a.mma_mode = s->desc.mma_mode;
  a.inv_B = 1.0f / (float)B_global; a.w_hat = s->desc.w_hat; a.w_ind = s->desc.w_ind;
  a.dt = (float)(m.T / s->N); a.q0 = (float)s->q0; a.R0 = (float)m.R0; a.S0 = (float)m.S0;
  a.alpha = (float)m.alpha; a.beta = (float)m.beta; a.jumpFactor = (float)m.jumpFactor; a.coeffOU = (float)m.coeffOU;
  a.A = (float)m.A; a.K = (float)m.K; a.pi = (float)m.pi; a.p0 = (float)m.p0; a.p1 = (float)m.p1; a.f0 = (float)m.f0;
  a.f1 = (float)m.f1; a.thetaR = (float)m.theta; a.C = (float)m.C; a.h1 = (float)m.h1; a.h2 = (float)m.h2;
  a.sig0 = (float)m.sig0; a.sig = (float)m.sig; a.alphaTarget = (float)m.alphaTarget; a.coeffEqui = (float)m.coeffEqui;
  a.qaver = s->qaver; a.meanhq = s->meanhq;
  a.netA = s->netA; a.netB = s->netB; a.y0_off = s->y0_off; a.P = s->P;
  a.theta = theta;
  a.dW0 = s->curA; a.dW = s->curB; a.dN = s->curC;
  a.traj = s->trajX; a.sch = s->sch1; a.fin = s->fin;
  a.lpart = s->lpart; a.gpart = s->gpart;
}

// forward (+ backward) + partial reduction into out[0 .. 4 (+P)).  Returns the number of kernels launched.
int run_pass(fbsdej_solver* s, const float* theta, int B, int B_global, float* out, bool with_grad, float* trajY,
             float* trajZ, cudaEvent_t* ev = nullptr) {
  FB_REQUIRE(B > 0 && B_global >= B, "B must be > 0 and B_global >= B");
  FB_REQUIRE(s->rng || s->noiseB == B, "noise not set for this batch size: call fbsdej_solver_simulate / set_noise first");
  cudaStream_t st = s->ctx->stream;
  int grid_f, grid_b = 0;
  if (s->model == FBSDEJ_MODEL_MFG) {
    MFGArgs a;
    fill_mfg_args(s, theta, B, B_global, a);
    const int ntiles = (B + kThreads - 1) / kThreads;
    const bool tc = a.mma_mode == 1;       // tcgen05 kernels: one CTA (two roles x 128 threads) per SM
    grid_f = std::min(ntiles, s->ctx->sms * (tc ? 1 : mfg_blocks_per_sm(s->HP, a, false)));
    if (with_grad) grid_b = std::min(ntiles, s->ctx->sms * (tc ? 1 : mfg_blocks_per_sm(s->HP, a, true)));
    if (ensure_grid(s, std::max(grid_f, grid_b))) return -2;
    a.lpart = s->lpart; a.gpart = s->gpart; a.trajY = trajY;
    { NvtxRange r("fbsdej:forward"); if (tc ? launch_mfg_tc(a, grid_f, false, st) : launch_mfg(s->HP, a, grid_f, false, st)) return -1; }
    if (ev) cudaEventRecord(ev[0], st);
    { NvtxRange r("fbsdej:adjoint"); if (with_grad && (tc ? launch_mfg_tc(a, grid_b, true, st) : launch_mfg(s->HP, a, grid_b, true, st))) return -1; }
    if (ev) cudaEventRecord(ev[1], st);
  } else {
    PricingArgs a;
    fill_pricing_args(s, theta, B, B_global, a);
    const int ppb = kThreads / a.G;
    const int ntiles = (B + ppb - 1) / ppb;
    // a.C CTAs (one cluster) per tile when the batch is small
    grid_f = a.C * std::min(ntiles, std::max(1, s->ctx->sms * pricing_blocks_per_sm(s->model, s->D, s->HP, a, false) / a.C));
    if (with_grad)
      grid_b = a.C * std::min(ntiles, std::max(1, s->ctx->sms * pricing_blocks_per_sm(s->model, s->D, s->HP, a, true) / a.C));
    if (a.mma_mode == 1 && !a.has_jump) {   // tcgen05 compensator-free kernels: four CTAs per SM, tiles of 4 or 3 warps
      if (!s->wimg && dev_alloc(&s->wimg, reg_tc_wimg_floats())) return -2;   // (first pass: never inside a graph capture)
      a.theta = theta;
      if (launch_reg_stage_operands(a, s->wimg, st)) return -1;             // theta -> TMA-ready operand images
      s->ctx->launches += 1;
      a.wimg_fwd = s->wimg; a.wimg_bwd = s->wimg + reg_tc_wimg_fwd_floats();
      a.tmap = make_tile_map(B, 4 * s->ctx->sms);
      grid_f = std::min(a.tmap.ntiles, 4 * s->ctx->sms);
      if (with_grad) grid_b = grid_f;
    }
    if (ensure_grid(s, std::max(grid_f, grid_b))) return -2;
    a.lpart = s->lpart; a.gpart = s->gpart; a.trajY = trajY; a.trajZ = trajZ;
    if (s->rng) {
      a.rng = 1; a.seed_lo = (uint32_t)s->rng->seed; a.seed_hi = (uint32_t)(s->rng->seed >> 32);
      a.iteration = s->rng->iteration; a.iter_ptr = s->rng->iter_ptr; a.path_offset = s->rng->path_offset;
      a.pois_thr = s->pois_thr; a.npois = s->npois;
      a.sqdt = (float)std::sqrt(s->mer.T / s->mer.N); a.muJ = (float)s->mer.muJ; a.sigJ = (float)s->mer.sigJ;
    }
    { NvtxRange r("fbsdej:forward"); if (launch_pricing(s->model, s->D, s->HP, a, grid_f, false, st)) return -1; }
    if (ev) cudaEventRecord(ev[0], st);
    { NvtxRange r("fbsdej:adjoint"); if (with_grad && launch_pricing(s->model, s->D, s->HP, a, grid_b, true, st)) return -1; }
    if (ev) cudaEventRecord(ev[1], st);
  }
  // loss partials come from the forward grid, gradient partials from the backward grid
  NvtxRange r_fin("fbsdej:reduce+adam");
  if (s->finish && with_grad) {
    const fbsdej_solver::Finish& f = *s->finish;
    XchgArgs x{};
    if (s->dp_step) {
      const size_t data_bytes = sizeof(float) * 2 * (size_t)s->dp.world * xchg_nstride(s->P);
      x.peer_data = s->dp.d_data; x.peer_flags = s->dp.d_flags;
      x.xctr = reinterpret_cast<uint32_t*>(s->dp.buf + data_bytes) + (size_t)s->dp.world * xchg_nblk(s->P);
      x.rank = s->dp.rank; x.world = s->dp.world; x.nstride = xchg_nstride(s->P); x.nblk = xchg_nblk(s->P);
      const char* to = getenv("FBSDEJ_DP_TIMEOUT_MS");              // default 30 s
      x.timeout_ns = (to && atof(to) > 0 ? (unsigned long long)(atof(to) * 1e6) : 30000000000ull);
    }
    if (launch_reduce_adam(s->lpart, grid_f, s->gpart, grid_b, s->P, out, f.theta, f.m, f.v, f.mask, f.lr, f.b1, f.b2, f.eps,
                           f.t_dev, f.iter_dev, f.loss_dst, s->step_ctr, s->step_ctr + 1, st, s->dp_step ? &x : nullptr))
      return -2;
  } else if (launch_reduce_partials(s->lpart, grid_f, s->gpart, grid_b, s->P, out, with_grad, st)) {
    return -2;
  }
  s->ctx->launches += with_grad ? 3 : 2;
  return 0;
}

int do_simulate(fbsdej_solver* s, uint64_t seed, uint32_t iteration, const uint32_t* iter_ptr, uint32_t path_offset,
                int B, cudaEvent_t* ev = nullptr) {
  if (ensure_capacity(s, B)) return -2;
  NvtxRange r_sim("fbsdej:simulate");
  cudaStream_t st = s->ctx->stream;
  const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
  if (s->model == FBSDEJ_MODEL_MERTON) {
    SimMertonArgs a{};
    a.B = B; a.N = s->N; a.D = s->D; a.seed_lo = lo; a.seed_hi = hi; a.iteration = iteration; a.iter_ptr = iter_ptr;
    a.path_offset = path_offset; a.stream = STREAM_PATH;
    a.sqdt = (float)std::sqrt(s->mer.T / s->mer.N); a.muJ = (float)s->mer.muJ; a.sigJ = (float)s->mer.sigJ;
    a.pois_thr = s->pois_thr; a.npois = s->npois; a.dW = s->nA; a.J = s->nB;
    if (launch_sim_merton(a, st)) return -2;
    if (ev) cudaEventRecord(*ev, st);
    s->ctx->launches += 1;
    if (s->has_jump) {
      SimMertonArgs c = a;
      c.B = s->M; c.path_offset = 0; c.stream = STREAM_JMC; c.dW = nullptr; c.J = s->jmc_raw;
      if (launch_sim_merton(c, st)) return -2;
      if (launch_compact_jmc(s->jmc_raw, s->jmc, s->jmc_nnz, s->jmc_n0, s->N, s->D, s->M, 1, st)) return -2;
      s->ctx->launches += 2;
    }
    s->curA = s->nA; s->curB = s->nB; s->curC = nullptr;
  } else if (s->model == FBSDEJ_MODEL_VG) {
    SimVGArgs a{};
    a.B = B; a.N = s->N; a.seed_lo = lo; a.seed_hi = hi; a.iteration = iteration; a.iter_ptr = iter_ptr;
    a.path_offset = path_offset; a.stream = STREAM_PATH;
    a.shape = (float)((s->vg.T / s->vg.N) / s->vg.kappa); a.scale = (float)s->vg.kappa;
    a.theta = (float)s->vg.theta; a.sigJ = (float)s->vg.sigJ; a.J = s->nB;
    if (launch_sim_vg(a, st)) return -2;
    if (ev) cudaEventRecord(*ev, st);
    s->ctx->launches += 1;
    if (s->has_jump) {
      SimVGArgs c = a;
      c.B = s->M; c.path_offset = 0; c.stream = STREAM_JMC; c.J = s->jmc_raw;
      if (launch_sim_vg(c, st)) return -2;
      if (launch_compact_jmc(s->jmc_raw, s->jmc, s->jmc_nnz, s->jmc_n0, s->N, 1, s->M, 1, st)) return -2;
      s->ctx->launches += 2;
    }
    s->curA = nullptr; s->curB = s->nB; s->curC = nullptr;
  } else {
    SimMFGArgs a{};
    a.B = B; a.N = s->N; a.seed_lo = lo; a.seed_hi = hi; a.iteration = iteration; a.iter_ptr = iter_ptr;
    a.path_offset = path_offset; a.stream = STREAM_MFG;
    a.dt = (float)(s->mfg.T / s->N); a.sqdt = (float)std::sqrt(s->mfg.T / s->N); a.q0 = (float)s->q0;
    a.alpha = (float)s->mfg.alpha; a.beta = (float)s->mfg.beta; a.jumpFactor = (float)s->mfg.jumpFactor;
    a.coeffOU = (float)s->mfg.coeffOU; a.sig0 = (float)s->mfg.sig0; a.stochastic = s->mfg.stochastic_jumps;
    a.qaver = s->qaver; a.dW0 = s->nA; a.dW = s->nB; a.dN = s->nC;
    if (launch_sim_mfg(a, st)) return -2;
    if (ev) cudaEventRecord(*ev, st);
    s->ctx->launches += 1;
    s->curA = s->nA; s->curB = s->nB; s->curC = s->nC;
  }
  s->noiseB = B;
  return 0;
}

// One training-path pass on fresh increments: simulate -> forward -> adjoint -> reduce.  The tcgen05 Merton solvers skip the
// simulation kernel: their forward sweep draws the same Philox increments in registers (bit-identical to do_simulate).
bool fused_rng(const fbsdej_solver* s) {
  return s->desc.mma_mode == 1 && s->model == FBSDEJ_MODEL_MERTON && !s->has_jump && !getenv("FBSDEJ_NO_FUSED_RNG");
}
int step_pass(fbsdej_solver* s, const float* theta, uint64_t seed, uint32_t iteration, const uint32_t* iter_ptr,
              uint32_t path_offset, int B, int B_global, float* out, cudaEvent_t* ev = nullptr) {
  if (!fused_rng(s)) {
    if (do_simulate(s, seed, iteration, iter_ptr, path_offset, B, ev ? &ev[1] : nullptr)) return -2;
    if (ev) cudaEventRecord(ev[2], s->ctx->stream);
    return run_pass(s, theta, B, B_global, out, true, nullptr, nullptr, ev ? &ev[3] : nullptr);
  }
  if (ensure_capacity(s, B)) return -2;
  if (ev) { cudaEventRecord(ev[1], s->ctx->stream); cudaEventRecord(ev[2], s->ctx->stream); }
  const fbsdej_solver::Rng rng{seed, iteration, iter_ptr, path_offset};
  s->rng = &rng;
  const int rc = run_pass(s, theta, B, B_global, out, true, nullptr, nullptr, ev ? &ev[3] : nullptr);
  s->rng = nullptr;
  s->noiseB = 0;                       // no increments were materialised
  return rc;
}

int build_merton_tables(fbsdej_solver* s) {
  const fbsdej_merton_params& m = s->mer;
  const int N = m.N, L = m.limit, d = m.d;
  const double dt = m.T / N;
  const double sigA = m.sig / std::sqrt((double)d), lamA = m.lam * d, muJA = m.muJ / d, sigJA = m.sigJ / d;
  const double kap = std::exp(m.muJ + 0.5 * m.sigJ * m.sigJ) - 1.0;
  const double kapA = std::exp(muJA + 0.5 * sigJA * sigJA) - 1.0;
  const double qA = d == 1 ? 0.0 : (0.5 * m.sig * m.sig + m.lam * kap - 0.5 * sigA * sigA - lamA * kapA);
  std::vector<float4> tA((size_t)N * L);
  std::vector<float> tK((size_t)N * L), qd(N);
  std::vector<int2> rg(N);
  for (int i = 0; i < N; ++i) {
    const double tau = m.T - i * dt, sq = std::sqrt(tau);
    const double lam2 = lamA * std::exp(muJA + 0.5 * sigJA * sigJA) * tau;
    qd[i] = (float)std::exp(-qA * tau);
    double wmax = 0.0;
    std::vector<double> w(L);
    for (int n = 0; n < L; ++n) {
      w[n] = std::exp(-lam2 + n * std::log(lam2) - std::lgamma(n + 1.0));
      wmax = std::max(wmax, w[n]);
    }
    int nlo = L, nhi = 0;
    for (int n = 0; n < L; ++n) {
      const double rn = m.r - lamA * kapA + n * (muJA + 0.5 * sigJA * sigJA) / tau;
      const double sn = std::sqrt(sigA * sigA + n * sigJA * sigJA / tau);
      const double c1 = 1.0 / (sn * sq), c2 = (rn + 0.5 * sn * sn) * tau / (sn * sq);
      tA[(size_t)i * L + n] = make_float4((float)c1, (float)c2, (float)(sn * sq), (float)w[n]);
      tK[(size_t)i * L + n] = (float)(w[n] * m.K * std::exp(-rn * tau));
      if (w[n] >= 1e-13 * wmax) { nlo = std::min(nlo, n); nhi = std::max(nhi, n + 1); }
    }
    rg[i] = make_int2(nlo, nhi);
  }
  cudaStream_t st = s->ctx->stream;
  if (dev_upload(&s->tabA, tA, st) || dev_upload(&s->tabK, tK, st) || dev_upload(&s->tab_range, rg, st) ||
      dev_upload(&s->qdisc, qd, st))
    return -2;
  s->limit = L;
  if (s->desc.price_table) {
    // Hermite table of sD(k) = sum_n w_n Phi(c1_n k + c2_n), sK(k) = sum_n wK_n Phi(c1_n k + c2_n - s_n) per step.
    // Spacing from the cubic-Hermite bound  h^4/384 * max|f|,  |Phi(x/s)| <= 0.55/s^4  (target 2e-8 absolute).
    const double kr = 1.6;
    std::vector<float4> nodes, meta(N);
    std::vector<int> offs(N);
    for (int i = 0; i < N; ++i) {
      const int2 r = rg[i];
      double acc = 0.0;
      for (int n = r.x; n < r.y; ++n) {
        const float4 c = tA[(size_t)i * L + n];
        acc += ((double)c.w + (double)tK[(size_t)i * L + n]) * std::pow((double)c.x, 4.0);   // w / s^4, c1 = 1/s
      }
      double h = std::pow(2e-8 * 384.0 / (0.55 * std::max(acc, 1e-30)), 0.25);
      h = std::min(h, 0.05);
      int nint = (int)std::ceil(2.0 * kr / h);
      nint = std::max(16, std::min(nint, 1 << 16));
      h = 2.0 * kr / nint;
      offs[i] = (int)nodes.size();
      meta[i] = make_float4((float)(-kr), (float)(1.0 / h), (float)h, (float)nint);
      for (int j = 0; j <= nint; ++j) {
        const double k = -kr + j * h;
        double sD = 0, dD = 0, sK = 0, dK = 0;
        for (int n = r.x; n < r.y; ++n) {
          const float4 c = tA[(size_t)i * L + n];     // fp32-rounded coefficients: the table matches the fp32 series
          const double wk = tK[(size_t)i * L + n];
          const double d1 = (double)c.x * k + (double)c.y, d2 = d1 - (double)c.z;
          sD += (double)c.w * 0.5 * std::erfc(-d1 * M_SQRT1_2);
          sK += wk * 0.5 * std::erfc(-d2 * M_SQRT1_2);
          dD += (double)c.w * (double)c.x * std::exp(-0.5 * d1 * d1) * 0.3989422804014327;
          dK += wk * (double)c.x * std::exp(-0.5 * d2 * d2) * 0.3989422804014327;
        }
        nodes.push_back(make_float4((float)sD, (float)dD, (float)sK, (float)dK));
      }
    }
    if (dev_upload(&s->atab, nodes, st) || dev_upload(&s->atab_meta, meta, st) || dev_upload(&s->atab_off, offs, st)) return -2;
    s->use_atab = 1;
  }
  // Poisson(lam dt) inversion table: thr[k] = floor(CDF(k) 2^32)
  const double mean = m.lam * dt;
  std::vector<uint32_t> thr;
  double p = std::exp(-mean), cdf = p;
  for (int k = 0; k < 64; ++k) {
    const double v = std::floor(cdf * 4294967296.0);
    if (v >= 4294967295.0) break;
    thr.push_back((uint32_t)v);
    p *= mean / (k + 1);
    cdf += p;
  }
  s->npois = (int)thr.size();
  if (thr.empty()) thr.push_back(0xffffffffu);
  if (dev_upload(&s->pois_thr, thr, st)) return -2;
  // drift of log X per step, formed the way the reference forms it in fp32 (pricingModels.py:54)
  const float e = std::exp((float)m.muJ + (float)m.sigJ * (float)m.sigJ * 0.5f);
  s->drift_dt = ((float)m.r - 0.5f * (float)m.sig * (float)m.sig - (float)m.lam * (e - 1.0f)) * (float)dt;
  return 0;
}

int upload_vg_table(fbsdej_solver* s, const double* coef, int nint, double k0, double h) {
  const int N = s->N;
  std::vector<float4> c((size_t)N * nint);
  for (size_t q = 0; q < c.size(); ++q)
    c[q] = make_float4((float)coef[4 * q], (float)coef[4 * q + 1], (float)coef[4 * q + 2], (float)coef[4 * q + 3]);
  std::vector<float> sc(N);
  for (int i = 0; i < N; ++i) sc[i] = (float)(std::exp(-s->vg.r * (s->vg.T - i * (s->vg.T / N))) / M_PI);
  dev_free(s->vg_coef); dev_free(s->vg_scale);
  if (dev_upload(&s->vg_coef, c, s->ctx->stream) || dev_upload(&s->vg_scale, sc, s->ctx->stream)) return -2;
  s->vg_nint = nint; s->vg_k0 = (float)k0; s->vg_h = (float)h;
  return 0;
}

int build_mfg_tables(fbsdej_solver* s, const double* QAver) {
  const int N = s->N;
  const double dt = s->mfg.T / N, k = s->mfg.coeffOU;
  std::vector<float> q(N + 1), mq(N + 1);
  for (int i = 0; i <= N; ++i) {
    q[i] = (float)QAver[i];
    double acc = 0.0;
    for (int j = 0; j < i; ++j) acc += QAver[j] * std::exp(k * (j - i) * dt) * dt;
    mq[i] = (float)(i == 0 ? QAver[0] : std::exp(-k * i * dt) * QAver[0] + k * acc);   // MFGModel.py:67-68
  }
  if (dev_upload(&s->qaver, q, s->ctx->stream) || dev_upload(&s->meanhq, mq, s->ctx->stream)) return -2;
  return 0;
}

}  // namespace

// =============================================================================================================
extern "C" {

const char* fbsdej_last_error(void) { return g_err.c_str(); }
int fbsdej_version(void) { return 100; }

int fbsdej_ctx_create(int device, void* stream, fbsdej_ctx** out) {
  FB_REQUIRE(out != nullptr, "ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  FB_CUDA(cudaGetDeviceCount(&n));
  FB_REQUIRE(device >= 0 && device < n, "ctx_create: no such CUDA device");
  FB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FB_CUDA(cudaGetDeviceProperties(&prop, device));
  FB_REQUIRE(prop.major == 10, "fbsdej is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                                   std::to_string(prop.minor));
  fbsdej_ctx* c = new fbsdej_ctx();
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->sms = prop.multiProcessorCount;
  *out = c;
  return 0;
}
int fbsdej_ctx_destroy(fbsdej_ctx* ctx) {
  delete ctx;
  return 0;
}
int fbsdej_ctx_sync(fbsdej_ctx* ctx) {
  FB_REQUIRE(ctx, "ctx is NULL");
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int fbsdej_malloc(fbsdej_ctx* ctx, size_t bytes, void** out) {
  FB_REQUIRE(ctx && out, "malloc: NULL argument");
  FB_CUDA(cudaSetDevice(ctx->device));
  FB_CUDA(cudaMalloc(out, bytes ? bytes : 4));
  return 0;
}
int fbsdej_free(fbsdej_ctx* ctx, void* p) {
  FB_REQUIRE(ctx, "ctx is NULL");
  if (p) FB_CUDA(cudaFree(p));
  return 0;
}
int fbsdej_memcpy_h2d(fbsdej_ctx* ctx, void* dst, const void* src_host, size_t bytes) {
  FB_REQUIRE(ctx, "ctx is NULL");
  FB_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int fbsdej_memcpy_d2h(fbsdej_ctx* ctx, void* dst_host, const void* src, size_t bytes) {
  FB_REQUIRE(ctx, "ctx is NULL");
  FB_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
long long fbsdej_ctx_launch_count(const fbsdej_ctx* ctx) { return ctx ? ctx->launches : -1; }

int fbsdej_solver_destroy(fbsdej_solver* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->graph) cudaGraphExecDestroy(s->graph);
  free_path_buffers(s);
  dev_free(s->tabA); dev_free(s->tabK); dev_free(s->tab_range); dev_free(s->qdisc);
  dev_free(s->atab); dev_free(s->atab_meta); dev_free(s->atab_off);
  dev_free(s->vg_coef); dev_free(s->vg_scale); dev_free(s->pois_thr); dev_free(s->qaver); dev_free(s->meanhq);
  dev_free(s->jmc_raw); dev_free(s->jmc); dev_free(s->jmc_nnz); dev_free(s->jmc_n0);
  dev_free(s->lpart); dev_free(s->gpart); dev_free(s->out_dev); dev_free(s->step_ctr); dev_free(s->wimg);
  for (size_t r = 0; r < s->dp.peers.size(); ++r)
    if (s->dp.opened[r] && s->dp.peers[r]) cudaIpcCloseMemHandle(s->dp.peers[r]);
  dev_free(s->dp.buf); dev_free(s->dp.d_data); dev_free(s->dp.d_flags);
  delete s;
  return 0;
}

int fbsdej_solver_create(fbsdej_ctx* ctx, const fbsdej_solver_desc* desc, const fbsdej_merton_params* merton,
                         const fbsdej_vg_params* vg, const fbsdej_mfg_params* mfg, fbsdej_solver** out) {
  FB_REQUIRE(ctx && desc && out, "solver_create: NULL argument");
  *out = nullptr;
  FB_CUDA(cudaSetDevice(ctx->device));
  const int model = desc->model, scheme = desc->scheme;
  FB_REQUIRE(model >= 0 && model <= 2, "solver_create: unknown model");
  FB_REQUIRE(scheme >= FBSDEJ_GLOBAL && scheme <= FBSDEJ_MULTISTEPREG, "solver_create: unknown scheme");
  FB_REQUIRE((model == FBSDEJ_MODEL_MERTON) == (merton != nullptr) && (model == FBSDEJ_MODEL_VG) == (vg != nullptr) &&
                 (model == FBSDEJ_MODEL_MFG) == (mfg != nullptr),
             "solver_create: exactly the parameter block matching desc->model must be given");
  const bool reg = scheme == FBSDEJ_SUMLOCALREG || scheme == FBSDEJ_MULTISTEPREG;
  const bool one_net = scheme == FBSDEJ_MULTISTEP1 || scheme == FBSDEJ_SUMLOCAL1;
  FB_REQUIRE(!(model == FBSDEJ_MODEL_MFG && one_net), "solver_create: the MFG solvers have no one-network variant");
  std::unique_ptr<fbsdej_solver, int (*)(fbsdej_solver*)> s(new fbsdej_solver(), fbsdej_solver_destroy);
  std::memset(&s->key, 0, sizeof(s->key));
  s->ctx = ctx;
  s->desc = *desc;
  s->model = model;
  s->sch = scheme == FBSDEJ_GLOBAL ? SCH_GLOBAL
           : (scheme == FBSDEJ_MULTISTEP1 || scheme == FBSDEJ_MULTISTEP2 || scheme == FBSDEJ_MULTISTEPREG) ? SCH_MULTISTEP
                                                                                                              : SCH_SUMLOCAL;
  s->one_net = one_net ? 1 : 0;
  s->has_jump = (!reg && model != FBSDEJ_MODEL_MFG) ? 1 : 0;
  s->has_y = scheme == FBSDEJ_GLOBAL ? 0 : 1;
  int D = 1, N = 0;
  if (merton) { s->mer = *merton; D = merton->d; N = merton->N; FB_REQUIRE(D == 1 || D == 10, "Merton: compiled for d in {1, 10}"); FB_REQUIRE(merton->limit >= 1 && merton->limit <= 4096, "Merton: limit out of range"); }
  if (vg) { s->vg = *vg; N = vg->N; }
  if (mfg) {
    s->mfg = *mfg;
    FB_REQUIRE(mfg->QAver && mfg->nQ >= 2, "MFG: QAver must hold nQ >= 2 values");
    N = mfg->nQ - 1;
  }
  FB_REQUIRE(N >= 1 && N <= 100000, "number of time steps out of range");
  s->D = D; s->N = N;
  // ---- expected network shapes (SURVEY 8a note i) -----------------------------------------------------------
  const int nn = desc->n_nets;
  FB_REQUIRE(nn == (one_net ? 1 : 2), "solver_create: n_nets must be 1 for the one-network schemes, else 2");
  int ninA, ninB, noutA, noutB, ny0;
  if (model == FBSDEJ_MODEL_MFG) {
    ninA = 4; ninB = 6; ny0 = scheme == FBSDEJ_GLOBAL ? 2 : 0;
    noutA = reg ? 1 : (scheme == FBSDEJ_GLOBAL ? 2 : 3);
    noutB = reg ? 1 : (scheme == FBSDEJ_GLOBAL ? 3 : 4);
    s->has_z = reg ? 0 : 1;
  } else {
    ninA = 1 + D; ninB = 1 + 2 * D; noutB = 1; ny0 = scheme == FBSDEJ_GLOBAL ? 1 : 0;
    const bool brown = model == FBSDEJ_MODEL_MERTON;
    if (scheme == FBSDEJ_GLOBAL) { noutA = brown ? D : 1; s->zoff = 0; s->has_z = brown ? 1 : 0; s->use_netA = brown ? 1 : 0; }
    else if (reg) { noutA = 1; s->has_z = 0; }
    else { noutA = brown ? 1 + D : 1; s->zoff = 1; s->has_z = brown ? 1 : 0; }
    s->feat_mode = scheme == FBSDEJ_GLOBAL ? 0 : 1;
  }
  FB_REQUIRE(desc->n_y0 == ny0, "solver_create: n_y0 must be " + std::to_string(ny0) + " for this scheme");
  const int expect_in[2] = {ninA, ninB}, expect_out[2] = {noutA, noutB};
  int off = 0, HP = 24;
  NetRt* nets[2] = {&s->netA, &s->netB};
  for (int k = 0; k < nn; ++k) {
    const fbsdej_net_desc& nd = desc->nets[k];
    FB_REQUIRE(nd.L >= 1 && nd.L <= kMaxL, "the fused kernels support 1, 2 or 3 hidden layers (the reference's nbLayer; default 2)");
    FB_REQUIRE(nd.nin == expect_in[k], "net " + std::to_string(k) + ": nin must be " + std::to_string(expect_in[k]));
    FB_REQUIRE(nd.nout == expect_out[k], "net " + std::to_string(k) + ": nout must be " + std::to_string(expect_out[k]));
    FB_REQUIRE(nd.act == FBSDEJ_ACT_TANH || nd.act == FBSDEJ_ACT_RELU, "unknown activation");
    const int hp = padded_width(nd.H);
    FB_REQUIRE(nd.H >= 1 && hp > 0, "hidden width must be in [1, 35]");
    FB_REQUIRE(hp <= 32 || (reg && model != FBSDEJ_MODEL_MFG), "hidden width 32..35 is compiled for the compensator-free pricing solvers only (else <= 31)");
    HP = std::max(HP, hp);
    FB_REQUIRE(nd.nin + 1 <= hp && nd.nout <= NOP, "network too wide for the compiled tiles");
    FB_REQUIRE(nd.L < 3 || hp == 24, "three hidden layers need a hidden width <= 23 (weight-gradient blocks per CTA)");
    nets[k]->nin = nd.nin; nets[k]->H = nd.H; nets[k]->nout = nd.nout; nets[k]->act = nd.act; nets[k]->ext_off = off; nets[k]->L = nd.L;
    off += net_ext_params(*nets[k]);
  }
  if (one_net) s->netB = s->netA;
  s->HP = HP;
  s->y0_off = off;
  s->P = off + ny0;
  s->M = s->has_jump ? desc->M : 0;
  FB_REQUIRE(!s->has_jump || desc->M >= 1, "this scheme needs M >= 1 compensator samples");
  FB_REQUIRE(desc->mma_mode == 0 || desc->mma_mode == 1, "mma_mode must be 0 (FFMA) or 1 (tcgen05)");
  bool all_two = true;
  for (int k = 0; k < nn; ++k) all_two = all_two && desc->nets[k].L == 2;
  FB_REQUIRE(desc->mma_mode == 0 || all_two, "mma_mode = 1 (tcgen05) needs two hidden layers");
  // tcgen05 coverage: the compensator-free pricing solvers, the MFG solvers, and the jump evaluations (own jump + Monte-Carlo
  // compensator) of the jump schemes with a tanh network
  const fbsdej_net_desc& jn = desc->nets[one_net ? 0 : 1];
  const bool jtc_ok = s->has_jump && (s->D == 1 || s->D == 10) && HP == 24 && jn.H <= 22 && jn.act == FBSDEJ_ACT_TANH && s->P <= 24 * kThreads;
  FB_REQUIRE(desc->mma_mode == 0 ||
                 (model == FBSDEJ_MODEL_MFG ? (desc->nets[0].H <= 22 && desc->nets[1].H <= 22 && desc->nets[0].act == desc->nets[1].act)
                                            : ((reg && HP == 24 && desc->nets[0].H <= 22) || jtc_ok)),
             "mma_mode = 1 (tcgen05) is available for the SUMLOCALREG / MULTISTEPREG pricing solvers, for the MFG solvers and for "
             "the jump schemes with a tanh jump network; hidden width <= 22");
  cudaStream_t st = ctx->stream;
  if (model == FBSDEJ_MODEL_MERTON) {
    if (build_merton_tables(s.get())) return -2;
  } else if (model == FBSDEJ_MODEL_VG) {
    std::vector<double> coef;
    double k0, h;
    build_vg_table(*vg, 128, coef, k0, h);
    if (upload_vg_table(s.get(), coef.data(), 256, k0, h)) return -2;
    s->drift_dt = (float)((vg->r - (-std::log(1.0 - vg->theta * vg->kappa - vg->kappa / 2.0 * vg->sigJ * vg->sigJ) / vg->kappa)) *
                          (vg->T / vg->N));
  } else {
    if (build_mfg_tables(s.get(), mfg->QAver)) return -2;
    s->q0 = mfg->QAver[0];
    s->mfg.QAver = nullptr;   // the caller's array is not retained
  }
  if (s->has_jump) {
    const size_t n = (size_t)N * D * s->M;
    if (dev_alloc(&s->jmc_raw, n) || dev_alloc(&s->jmc, n) || dev_alloc(&s->jmc_nnz, (size_t)N) || dev_alloc(&s->jmc_n0, (size_t)N))
      return -2;
  }
  if (dev_alloc(&s->out_dev, (size_t)(kHeader + s->P)) || dev_alloc(&s->step_ctr, (size_t)2)) return -2;
  FB_CUDA(cudaMemsetAsync(s->step_ctr, 0, 2 * sizeof(uint32_t), st));
  FB_CUDA(cudaStreamSynchronize(st));
  *out = s.release();
  return 0;
}

int fbsdej_solver_nparams(const fbsdej_solver* s) { return s ? s->P : -1; }

int fbsdej_solver_set_weights(fbsdej_solver* s, float w_hat, float w_ind) {
  FB_REQUIRE(s, "set_weights: solver is NULL");
  if (s->desc.w_hat != w_hat || s->desc.w_ind != w_ind) {
    if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
    s->desc.w_hat = w_hat; s->desc.w_ind = w_ind;
  }
  return 0;
}

int fbsdej_solver_set_vg_table_host(fbsdej_solver* s, const double* coef, int n_int, double k0, double h) {
  FB_REQUIRE(s && coef && n_int >= 1 && h > 0, "set_vg_table: bad argument");
  FB_REQUIRE(s->model == FBSDEJ_MODEL_VG, "set_vg_table: not a VG solver");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  return upload_vg_table(s, coef, n_int, k0, h);
}

int fbsdej_solver_simulate(fbsdej_solver* s, uint64_t seed, uint32_t iteration, uint32_t path_offset, int B) {
  FB_REQUIRE(s && B > 0, "simulate: bad argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  return do_simulate(s, seed, iteration, nullptr, path_offset, B);
}

int fbsdej_solver_set_noise(fbsdej_solver* s, int B, const float* a, const float* b, const float* c) {
  FB_REQUIRE(s && B > 0, "set_noise: bad argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  if (ensure_capacity(s, B)) return -2;
  if (s->model == FBSDEJ_MODEL_MFG) {
    FB_REQUIRE(a && b && c, "set_noise (MFG): dW0, dW and dN are all required");
    s->curA = a; s->curB = b; s->curC = c;
  } else {
    FB_REQUIRE(b, "set_noise: J is required");
    FB_REQUIRE((s->model == FBSDEJ_MODEL_MERTON) == (a != nullptr), "set_noise: dW is required for Merton and must be NULL for VG");
    FB_REQUIRE(!s->has_jump || c, "set_noise: JMC is required for schemes with a compensator");
    s->curA = a; s->curB = b; s->curC = nullptr;
    if (s->has_jump) {
      if (launch_compact_jmc(c, s->jmc, s->jmc_nnz, s->jmc_n0, s->N, s->D, s->M, 1, s->ctx->stream)) return -2;
      s->ctx->launches += 1;
    }
  }
  s->noiseB = B;
  return 0;
}

int fbsdej_solver_set_noise_sparse_jumps(fbsdej_solver* s, int B, const float* dW, const uint32_t* jidx, const float* jval, int nnz,
                                         const float* jmc) {
  FB_REQUIRE(s && B > 0 && nnz >= 0 && (nnz == 0 || (jidx && jval)), "set_noise_sparse_jumps: bad argument");
  FB_REQUIRE(s->model != FBSDEJ_MODEL_MFG, "set_noise_sparse_jumps: pricing models only");
  FB_REQUIRE((s->model == FBSDEJ_MODEL_MERTON) == (dW != nullptr), "set_noise_sparse_jumps: dW is required for Merton and must be NULL for VG");
  FB_REQUIRE(!s->has_jump || jmc, "set_noise_sparse_jumps: JMC is required for schemes with a compensator");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  if (ensure_capacity(s, B)) return -2;
  // the dense jump planes are rebuilt in the solver's own increment buffer (the one fbsdej_solver_simulate fills)
  if (launch_scatter(s->nB, (size_t)s->N * s->D * B, jidx, jval, nnz, s->ctx->stream)) return -2;
  s->ctx->launches += 1;
  s->curA = dW; s->curB = s->nB; s->curC = nullptr;
  if (s->has_jump) {
    if (launch_compact_jmc(jmc, s->jmc, s->jmc_nnz, s->jmc_n0, s->N, s->D, s->M, 1, s->ctx->stream)) return -2;
    s->ctx->launches += 1;
  }
  s->noiseB = B;
  return 0;
}

int fbsdej_solver_get_noise(fbsdej_solver* s, const float** a, const float** b, const float** c, const int** jmc_nnz,
                            const int** jmc_n0) {
  FB_REQUIRE(s, "get_noise: solver is NULL");
  if (a) *a = s->curA;
  if (b) *b = s->curB;
  if (c) *c = s->model == FBSDEJ_MODEL_MFG ? s->curC : s->jmc;
  if (jmc_nnz) *jmc_nnz = s->jmc_nnz;
  if (jmc_n0) *jmc_n0 = s->jmc_n0;
  return 0;
}

int fbsdej_solver_loss(fbsdej_solver* s, const float* theta, int B, int B_global, float* out, float* trajX,
                       float* trajY, float* trajZ) {
  FB_REQUIRE(s && theta && out, "loss: NULL argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  const int rc = run_pass(s, theta, B, B_global, out, false, trajY, trajZ);
  if (rc) return rc;
  if (trajX) {
    if (s->model == FBSDEJ_MODEL_MFG) {
      // (hS, S) planes of the internal [N+1][5][B] state dump
      for (int i = 0; i <= s->N; ++i)
        FB_CUDA(cudaMemcpyAsync(trajX + (size_t)i * 2 * B, s->trajX + ((size_t)i * 5 + 3) * B, sizeof(float) * 2 * B,
                                cudaMemcpyDeviceToDevice, s->ctx->stream));
    } else if (s->desc.mma_mode == 1 && !s->has_jump) {
      if (launch_untile_traj(s->D, s->rec, s->recN, make_tile_map(B, 4 * s->ctx->sms), B, s->N, trajX, s->ctx->stream)) return -1;
    } else {
      FB_CUDA(cudaMemcpyAsync(trajX, s->trajX, sizeof(float) * (size_t)(s->N + 1) * s->D * B, cudaMemcpyDeviceToDevice,
                              s->ctx->stream));
    }
  }
  return 0;
}

int fbsdej_solver_mfg_states(fbsdej_solver* s, int B, float* out) {
  FB_REQUIRE(s && out && B > 0, "mfg_states: bad argument");
  FB_REQUIRE(s->model == FBSDEJ_MODEL_MFG, "mfg_states: not an MFG solver");
  FB_REQUIRE(s->trajX && B <= s->capB, "mfg_states: run fbsdej_solver_loss / fbsdej_solver_grad with this batch size first");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  FB_CUDA(cudaMemcpyAsync(out, s->trajX, sizeof(float) * (size_t)(s->N + 1) * 5 * B, cudaMemcpyDeviceToDevice, s->ctx->stream));
  return 0;
}

int fbsdej_solver_grad(fbsdej_solver* s, const float* theta, int B, int B_global, float* out) {
  FB_REQUIRE(s && theta && out, "grad: NULL argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  return run_pass(s, theta, B, B_global, out, true, nullptr, nullptr);
}

int fbsdej_adam_step(fbsdej_ctx* ctx, float* theta, float* m, float* v, const float* grad, const float* mask, int n,
                     float lr, float beta1, float beta2, float eps, int* t_dev) {
  FB_REQUIRE(ctx && theta && m && v && grad && t_dev && n > 0, "adam_step: bad argument");
  FB_CUDA(cudaSetDevice(ctx->device));
  if (launch_adam(theta, m, v, grad, mask, n, lr, beta1, beta2, eps, t_dev, ctx->stream)) return -2;
  ctx->launches += 2;
  return 0;
}

int fbsdej_solver_grad_step(fbsdej_solver* s, const float* theta, uint64_t seed, const uint32_t* iter_dev,
                            uint32_t path_offset, int B, int B_global, float* out) {
  FB_REQUIRE(s && theta && out && iter_dev, "grad_step: NULL argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  return step_pass(s, theta, seed, 0, iter_dev, path_offset, B, B_global, out);
}

int fbsdej_bump_u32(fbsdej_ctx* ctx, uint32_t* p) {
  FB_REQUIRE(ctx && p, "bump: NULL argument");
  if (launch_bump_u32(p, ctx->stream)) return -2;
  ctx->launches += 1;
  return 0;
}

namespace {
int train_steps_impl(fbsdej_solver* s, float* theta, float* m, float* v, const float* mask, int* t_dev, uint32_t* iter_dev,
                     uint64_t seed, int B, int B_global, uint32_t path_offset, bool dp, int n_steps, float lr, float beta1,
                     float beta2, float eps, float* loss_out) {
  FB_REQUIRE(s && theta && m && v && t_dev && iter_dev && B > 0 && n_steps >= 0, "train_steps: bad argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  cudaStream_t st = s->ctx->stream;
  FB_REQUIRE(st != nullptr, "train_steps: stream capture needs a non-default stream in the ctx");
  fbsdej_solver::Key k;
  std::memset(&k, 0, sizeof(k));
  k.theta = theta; k.m = m; k.v = v; k.mask = mask; k.t = t_dev; k.it = iter_dev; k.loss = loss_out;
  k.seed = seed; k.B = B; k.lr = lr; k.b1 = beta1; k.b2 = beta2; k.eps = eps;
  k.B_global = B_global; k.path_offset = path_offset; k.dp = dp ? 1 : 0;
  const bool same = s->graph && std::memcmp(&k, &s->key, sizeof(k)) == 0;
  if (!same) {
    if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
    // one eager step sizes every buffer (allocation is illegal during capture); it is a real training step
    if (ensure_capacity(s, B)) return -2;
  }
  FB_CUDA(cudaMemsetAsync(s->step_ctr, 0, sizeof(uint32_t), st));
  int done = 0;
  const fbsdej_solver::Finish fin{theta, m, v, mask, lr, beta1, beta2, eps, t_dev, iter_dev, loss_out};
  auto one_step = [&]() -> int {       // simulate, forward, adjoint, [reduce + Adam + counters + loss record]: 4 launches
    s->finish = &fin;
    s->dp_step = dp;
    const int rc = step_pass(s, theta, seed, 0, iter_dev, path_offset, B, B_global, s->out_dev);
    s->finish = nullptr;
    s->dp_step = false;
    return rc ? -2 : 0;
  };
  if (!same && n_steps > 0) {
    if (one_step()) return -2;
    done = 1;
    const long long before = s->ctx->launches;
    cudaGraph_t g = nullptr;
    FB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = one_step();
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (rc || ce != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      if (!rc) set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
      return -2;
    }
    s->launches_per_step = s->ctx->launches - before;
    s->ctx->launches = before;
    FB_CUDA(cudaGraphInstantiate(&s->graph, g, 0));
    cudaGraphDestroy(g);
    s->key = k;
  }
  for (; done < n_steps; ++done) {
    FB_CUDA(cudaGraphLaunch(s->graph, st));
    s->ctx->launches += s->launches_per_step;
  }
  return 0;
}
}  // namespace

int fbsdej_solver_train_steps(fbsdej_solver* s, float* theta, float* m, float* v, const float* mask, int* t_dev,
                              uint32_t* iter_dev, uint64_t seed, int B, int n_steps, float lr, float beta1,
                              float beta2, float eps, float* loss_out) {
  return train_steps_impl(s, theta, m, v, mask, t_dev, iter_dev, seed, B, B, 0u, false, n_steps, lr, beta1, beta2, eps, loss_out);
}

// ---- data-parallel training steps without a host-driven collective --------------------------------------------------
// Every rank (one process per GPU, or several solvers of one process in the tests) owns an exchange buffer; dp_init
// allocates it and returns its CUDA IPC handle, the caller gathers the handles of all ranks (torch.distributed, MPI, ...)
// and hands them to dp_connect, which maps the peers' buffers (NVLink peer access) - or takes raw device pointers for
// ranks that live in the same process.  train_steps_dp is train_steps on this rank's shard [path_offset, path_offset + B)
// of a global batch: the finishing kernel exchanges the [loss | gradient] vector through the peers' buffers and every
// rank applies the same Adam update - one CUDA graph per step, no NCCL call, no host synchronisation between the ranks.
int fbsdej_solver_dp_init(fbsdej_solver* s, int rank, int world, unsigned char* handle64) {
  FB_REQUIRE(s && handle64 && world >= 1 && world <= 32 && rank >= 0 && rank < world, "dp_init: bad argument");
  FB_REQUIRE(!s->dp.buf, "dp_init: already initialised");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  const size_t bytes = xchg_bytes(s->P, world);
  if (dev_alloc(&s->dp.buf, bytes)) return -2;
  FB_CUDA(cudaMemset(s->dp.buf, 0, bytes));
  s->dp.rank = rank; s->dp.world = world;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  FB_CUDA(cudaIpcGetMemHandle(&h, s->dp.buf));
  std::memcpy(handle64, &h, 64);
  return 0;
}
int fbsdej_solver_dp_buffer(fbsdej_solver* s, void** ptr) {
  FB_REQUIRE(s && ptr && s->dp.buf, "dp_buffer: dp_init first");
  *ptr = s->dp.buf;
  return 0;
}
int fbsdej_solver_dp_connect(fbsdej_solver* s, const unsigned char* handles, void* const* raw_ptrs) {
  FB_REQUIRE(s && s->dp.buf && (handles || raw_ptrs), "dp_connect: dp_init first; handles or raw pointers needed");
  FB_REQUIRE(!s->dp.connected, "dp_connect: already connected");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  const int W = s->dp.world;
  s->dp.peers.assign(W, nullptr); s->dp.opened.assign(W, 0);
  for (int r = 0; r < W; ++r) {
    if (r == s->dp.rank) { s->dp.peers[r] = s->dp.buf; continue; }
    if (raw_ptrs && raw_ptrs[r]) { s->dp.peers[r] = raw_ptrs[r]; continue; }
    FB_REQUIRE(handles != nullptr, "dp_connect: no handle for rank " + std::to_string(r));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    FB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->dp.peers[r] = p; s->dp.opened[r] = 1;
  }
  const size_t data_bytes = sizeof(float) * 2 * (size_t)W * xchg_nstride(s->P);
  std::vector<float*> hd(W); std::vector<uint32_t*> hf(W);
  for (int r = 0; r < W; ++r) {
    hd[r] = reinterpret_cast<float*>(s->dp.peers[r]);
    hf[r] = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(s->dp.peers[r]) + data_bytes);
  }
  if (dev_alloc(&s->dp.d_data, (size_t)W) || dev_alloc(&s->dp.d_flags, (size_t)W)) return -2;
  FB_CUDA(cudaMemcpy(s->dp.d_data, hd.data(), sizeof(float*) * W, cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(s->dp.d_flags, hf.data(), sizeof(uint32_t*) * W, cudaMemcpyHostToDevice));
  s->dp.connected = true;
  return 0;
}
// Synchronises the ctx stream and reads this rank's error word: 0, or the stamp of the first exchange that timed out (the
// step was void on this rank: no parameter / Adam / counter update).  Returns -3 with a message in that case.
int fbsdej_solver_dp_check(fbsdej_solver* s) {
  FB_REQUIRE(s && s->dp.buf, "dp_check: dp_init first");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  FB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  const size_t data_bytes = sizeof(float) * 2 * (size_t)s->dp.world * xchg_nstride(s->P);
  const uint32_t* ctr = reinterpret_cast<const uint32_t*>(s->dp.buf + data_bytes) + (size_t)s->dp.world * xchg_nblk(s->P);
  uint32_t w[2] = {0, 0};
  FB_CUDA(cudaMemcpy(w, ctr, sizeof(w), cudaMemcpyDeviceToHost));
  if (w[1] != 0u) {
    set_error("data-parallel exchange timed out at step stamp " + std::to_string(w[1]) + " on rank " + std::to_string(s->dp.rank) +
              " of " + std::to_string(s->dp.world) + ": a peer did not deliver its [loss | gradient] vector; the step was not applied");
    return -3;
  }
  return 0;
}
int fbsdej_solver_train_steps_dp(fbsdej_solver* s, float* theta, float* m, float* v, const float* mask, int* t_dev,
                                 uint32_t* iter_dev, uint64_t seed, int B, int B_global, uint32_t path_offset, int n_steps,
                                 float lr, float beta1, float beta2, float eps, float* loss_out) {
  FB_REQUIRE(s && s->dp.connected, "train_steps_dp: fbsdej_solver_dp_init / dp_connect first");
  FB_REQUIRE(B_global >= B, "train_steps_dp: B_global must be >= B");
  return train_steps_impl(s, theta, m, v, mask, t_dev, iter_dev, seed, B, B_global, path_offset, true, n_steps, lr, beta1, beta2,
                          eps, loss_out);
}

int fbsdej_solver_profile(fbsdej_solver* s, const float* theta, uint64_t seed, int B, int reps, float* ms_host) {
  FB_REQUIRE(s && theta && ms_host && B > 0 && reps > 0, "profile: bad argument");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  cudaStream_t st = s->ctx->stream;
  cudaEvent_t ev[6];
  for (auto& e : ev) FB_CUDA(cudaEventCreate(&e));
  for (int k = 0; k < 5; ++k) ms_host[k] = 0.0f;
  int rc = 0;
  for (int r = -1; r < reps && !rc; ++r) {          // r == -1: untimed warm-up (also sizes the buffers)
    FB_CUDA(cudaEventRecord(ev[0], st));
    rc = step_pass(s, theta, seed, (uint32_t)(r + 1), nullptr, 0, B, B, s->out_dev, ev);
    if (rc) break;
    FB_CUDA(cudaEventRecord(ev[5], st));
    FB_CUDA(cudaStreamSynchronize(st));
    if (r < 0) continue;
    for (int k = 0; k < 5; ++k) {
      float ms = 0.0f;
      FB_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      ms_host[k] += ms / (float)reps;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int fbsdej_solver_net_forward(fbsdej_solver* s, const float* theta, int net_index, const float* x, int rows, float* y) {
  FB_REQUIRE(s && theta && x && y && rows > 0, "net_forward: bad argument");
  FB_REQUIRE(net_index == 0 || (net_index == 1 && !s->one_net), "net_forward: no such net");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  const NetRt& n = net_index == 0 ? s->netA : s->netB;
  if (launch_net_forward(theta + n.ext_off, n.nin, n.H, n.L, n.nout, n.act, x, rows, y, s->ctx->stream)) return -2;
  s->ctx->launches += 1;
  return 0;
}

int fbsdej_net_forward(fbsdej_ctx* ctx, const float* theta_net, int nin, int H, int L, int nout, int act, const float* x,
                       int rows, float* y) {
  FB_REQUIRE(ctx && theta_net && x && y && rows > 0, "net_forward: bad argument");
  FB_CUDA(cudaSetDevice(ctx->device));
  if (launch_net_forward(theta_net, nin, H, L, nout, act, x, rows, y, ctx->stream)) return -2;
  ctx->launches += 1;
  return 0;
}

int fbsdej_solver_price(fbsdej_solver* s, int iStep, const float* X, int n, float* out) {
  FB_REQUIRE(s && X && out && n > 0 && iStep >= 0, "price: bad argument");
  FB_REQUIRE(s->model != FBSDEJ_MODEL_MFG, "price: the MFG model has no closed-form price");
  FB_CUDA(cudaSetDevice(s->ctx->device));
  PricingArgs a;
  fill_pricing_args(s, nullptr, 1, 1, a);
  if (launch_price(s->model, s->D, a, iStep, X, n, out, s->ctx->stream)) return -2;
  s->ctx->launches += 1;
  return 0;
}


int fbsdej_transpose_nbd_to_ndb(fbsdej_ctx* ctx, const float* src, float* dst, int N, int B, int d) {
  FB_REQUIRE(ctx && src && dst, "transpose: NULL argument");
  if (launch_transpose(src, dst, N, B, d, true, ctx->stream)) return -2;
  ctx->launches += 1;
  return 0;
}
int fbsdej_transpose_ndb_to_nbd(fbsdej_ctx* ctx, const float* src, float* dst, int N, int B, int d) {
  FB_REQUIRE(ctx && src && dst, "transpose: NULL argument");
  if (launch_transpose(src, dst, N, B, d, false, ctx->stream)) return -2;
  ctx->launches += 1;
  return 0;
}

}  // extern "C"

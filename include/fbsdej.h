/* fbsdej.h - C-ABI of the B200 (sm_100a) deep-FBSDE-with-jumps training library.
 *
 * The reference (ZakariaBensaid/DeepFBSDEJSolvers) has no FFI: its boundary is the Python class API
 * (SURVEY.md 8b).  This header is the boundary the Python shim in deepfbsdejsolvers_b200/ binds with
 * ctypes; each entry point names the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; fbsdej_last_error() gives the message
 *    (thread-local).  No C++ exceptions cross the boundary.
 *  - fbsdej_ctx = (device, stream, workspace).  NOT thread-safe per ctx; one ctx per GPU / rank.
 *  - all tensor arguments are caller-owned DEVICE pointers to contiguous fp32 unless stated; calls are
 *    asynchronous on the ctx stream.  Host-pointer convenience calls say so in their name (_host).
 *  - path tensors are time-major component planes: dW/J [N][d][B], JMC [N][d][M], trajectories
 *    X [N+1][d][B], Y [N+1][B], Z [N][d][B]  (a step slice is coalesced over paths).
 *  - flat parameter vector: per net, per layer W[in][out] row-major then b[out]; nets concatenated
 *    (net 0 = UZ/U or model_hat, net 1 = Gam or model); trainable Y0 scalars last.
 *  - result vector `out` of loss/grad calls: out[0] = loss, out[1] = loss_a (MFG hat player),
 *    out[2] = loss_b (MFG individual player), out[3] = reserved, out[4 .. 4+P) = d loss / d theta.
 *    With B_global > B (data parallel) each rank returns its share of the global mean; summing the
 *    vectors over ranks (one all-reduce) gives the exact global loss and gradient.
 */
#ifndef FBSDEJ_H
#define FBSDEJ_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#define FBSDEJ_API __attribute__((visibility("default")))
#else
#define FBSDEJ_API
#endif

typedef struct fbsdej_ctx fbsdej_ctx;
typedef struct fbsdej_solver fbsdej_solver;

#define FBSDEJ_OUT_HEADER 4

enum { FBSDEJ_MODEL_MERTON = 0, FBSDEJ_MODEL_VG = 1, FBSDEJ_MODEL_MFG = 2 };
/* loss graphs; pricing: SolversJumpDiff.py / SolversPureJump.py classes, MFG: MFGSolvers.py classes */
enum {
  FBSDEJ_GLOBAL = 0,        /* SolverGlobalFBSDE            (JumpDiff :17-73, PureJump :17-72, MFG :17-116)   */
  FBSDEJ_MULTISTEP1 = 1,    /* SolverMultiStepFBSDE1        (JumpDiff :75-149, PureJump :74-141)              */
  FBSDEJ_MULTISTEP2 = 2,    /* SolverMultiStepFBSDE2 / MFG SolverMultiStepFBSDE (:151-224 / MFG :180-294)     */
  FBSDEJ_SUMLOCAL1 = 3,     /* SolverSumLocalFBSDE1         (JumpDiff :226-303, PureJump :210-280)            */
  FBSDEJ_SUMLOCAL2 = 4,     /* SolverSumLocalFBSDE2 / MFG SolverSumLocalFBSDE (:305-381 / MFG :321-434)       */
  FBSDEJ_SUMLOCALREG = 5,   /* SolverGlobalSumLocalReg      (JumpDiff :385-445, PureJump :355-414, MFG :463-579) */
  FBSDEJ_MULTISTEPREG = 6   /* SolverGlobalMultiStepReg     (JumpDiff :453-513, PureJump :422-482, MFG :608-725) */
};
enum { FBSDEJ_ACT_TANH = 0, FBSDEJ_ACT_RELU = 1 };

/* MertonJumpModel(T,N,r,muJ,sigmaJ,sigma,lam,K,x0,func,limit) pricingModels.py:11-24; func = aLin*|x|
 * (mainMerton.py:60-61).  d > 1: independent assets, geometric-basket payoff. */
typedef struct {
  double T, r, muJ, sigJ, sig, lam, K, x0, aLin;
  int N, limit, d;
} fbsdej_merton_params;

/* VGmodel(T,N,r,theta,kappa,sigmaJ,K,x0,func) pricingModels.py:131-141.  The Lewis-FFT price table of
 * pricingModels.py:156-179 is built inside the library (float64 FFT + not-a-knot cubic spline). */
typedef struct {
  double T, r, theta, kappa, sigJ, K, x0, aLin;
  int N;
} fbsdej_vg_params;

/* ModelCoupledFBSDE(...) MFGModel.py:5-31.  QAver: HOST pointer to nQ = N+1 doubles (copied). */
typedef struct {
  double T, R0, jumpFactor, alpha, beta, coeffOU, A, K, pi, p0, p1, f0, f1, theta, C, S0, h1, h2, sig0, sig,
      alphaTarget, coeffEqui;
  int stochastic_jumps; /* jumpModel == 'stochastic' */
  int nQ;
  const double* QAver;
} fbsdej_mfg_params;

/* Net(bY0, ndimOut, nbNeurons, activation) Networks.py:6-15: nin -> H x L -> nout. */
typedef struct {
  int nin, nout, H, L, act;
} fbsdej_net_desc;

typedef struct {
  int model;            /* FBSDEJ_MODEL_* */
  int scheme;           /* FBSDEJ_*       */
  int n_nets;           /* 1 or 2         */
  fbsdej_net_desc nets[2];
  int n_y0;             /* trainable scalars after the nets (Global: 1 pricing, 2 MFG) */
  int M;                /* compensator samples (reference: 5000, SolversJumpDiff.py:34); 0 for *Reg / MFG */
  int stale_time;       /* 1 = reference behaviour of the SumLocal graphs */
  float w_hat, w_ind;   /* MFG objective = w_hat*loss_hat + w_ind*loss_ind (couplage ON: 1,1) */
  int price_table;      /* Merton: 0 = sum the series of pricingModels.py:40-49 term by term at every path-step,
                           1 = evaluate it through a per-step cubic-Hermite table in log-moneyness (abs. error < 5e-8,
                           the same idea as the reference's own spline of the VG price, pricingModels.py:170-178) */
  int mma_mode;         /* 0 = every layer in fp32 FFMA; 1 = every matrix product of the network on tcgen05 tensor cores
                           (forward: both layers 3xTF32 with the A operands in TMEM; adjoint: six bf16x3 GEMMs per step,
                           weight-gradient accumulators resident in TMEM).  Available for the compensator-free solvers
                           (SUMLOCALREG / MULTISTEPREG) of the pricing models and for all five MFG solvers, H <= 22; for the
                           jump schemes (tanh) it moves the jump evaluations - the path's own jump and the
                           Monte-Carlo compensator rows - onto tcgen05 (the (U, Z) network stays fp32 FFMA). */
} fbsdej_solver_desc;

FBSDEJ_API const char* fbsdej_last_error(void);
FBSDEJ_API int fbsdej_version(void);

/* stream: a cudaStream_t (may be NULL = default stream) owned by the caller. */
FBSDEJ_API int fbsdej_ctx_create(int device, void* stream, fbsdej_ctx** out);
FBSDEJ_API int fbsdej_ctx_destroy(fbsdej_ctx* ctx);
FBSDEJ_API int fbsdej_ctx_sync(fbsdej_ctx* ctx);
FBSDEJ_API int fbsdej_malloc(fbsdej_ctx* ctx, size_t bytes, void** out);
FBSDEJ_API int fbsdej_free(fbsdej_ctx* ctx, void* p);
FBSDEJ_API int fbsdej_memcpy_h2d(fbsdej_ctx* ctx, void* dst, const void* src_host, size_t bytes);
FBSDEJ_API int fbsdej_memcpy_d2h(fbsdej_ctx* ctx, void* dst_host, const void* src, size_t bytes);

/* Solver objects.  Exactly one of merton/vg/mfg must be non-NULL and match desc->model. */
FBSDEJ_API int fbsdej_solver_create(fbsdej_ctx* ctx, const fbsdej_solver_desc* desc, const fbsdej_merton_params* merton,
                         const fbsdej_vg_params* vg, const fbsdej_mfg_params* mfg, fbsdej_solver** out);
FBSDEJ_API int fbsdej_solver_destroy(fbsdej_solver* s);
FBSDEJ_API int fbsdej_solver_nparams(const fbsdej_solver* s);
/* MFG objective weights (couplage OFF trains the hat player, then the individual player: MFGSolvers.py:92-115). */
FBSDEJ_API int fbsdej_solver_set_weights(fbsdej_solver* s, float w_hat, float w_ind);
/* Optional: replace the VG price table (host float64 [N][n_int][4] cubic coefficients, knots k0 + j*h). */
FBSDEJ_API int fbsdej_solver_set_vg_table_host(fbsdej_solver* s, const double* coef, int n_int, double k0, double h);

/* Noise.  Replaces tf.random.normal / mathModel.jumps / mathModel.dN draws (SolversJumpDiff.py:30-34,
 * pricingModels.py:57-61,188-191, MFGModel.py:47-54, MFGSolvers.py:35-38).
 * simulate: counter-based Philox4x32-10, counter = (path_offset + b, step, iteration, stream); identical
 *           noise for a path whatever the sharding.  JMC is generated identically on every rank.
 * set_noise: injection hook for parity tests (device pointers; NULL where the model has no such input).
 *   pricing: dW [N][d][B] (Merton), J [N][d][B], JMC [N][d][M];  MFG: a = dW0, b = dW, c = dN, all [N][B]. */
FBSDEJ_API int fbsdej_solver_simulate(fbsdej_solver* s, uint64_t seed, uint32_t iteration, uint32_t path_offset, int B);
FBSDEJ_API int fbsdej_solver_set_noise(fbsdej_solver* s, int B, const float* a, const float* b, const float* c);
/* The same for the pricing models with the jump planes given as their non-zero entries (a compound-Poisson increment is
 * exactly 0 wherever no jump fell into the step - 97 % of the entries at lam dt = 0.03): J[jidx[i]] = jval[i], jidx = flat
 * index into [N][d][B], everything else 0.  Lossless; halves the host-to-device bytes of an injected step.  dW / jmc as in
 * set_noise (device pointers). */
FBSDEJ_API int fbsdej_solver_set_noise_sparse_jumps(fbsdej_solver* s, int B, const float* dW, const uint32_t* jidx, const float* jval,
                                         int nnz, const float* jmc);
/* Device pointers of the solver's current noise tensors (for dumps / statistics tests). */
FBSDEJ_API int fbsdej_solver_get_noise(fbsdej_solver* s, const float** a, const float** b, const float** c, const int** jmc_nnz,
                            const int** jmc_n0);

/* Forward only (validation loss + trajectory dump): optimizeBSDE / regressOptim without the tape.
 * trajectories may be NULL. pricing: X [N+1][d][B], Y [N+1][B], Z [N][d][B].
 * MFG: X = (hS,S) [N+1][2][B], Y = (hY,Y) [N+1][2][B], Z ignored. out: >= 4 floats. */
FBSDEJ_API int fbsdej_solver_loss(fbsdej_solver* s, const float* theta, int B, int B_global, float* out, float* trajX,
                       float* trajY, float* trajZ);

/* MFG: the full state dump of the last fbsdej_solver_loss / fbsdej_solver_grad call, out [N+1][5][B] with planes
 * (hQ, Q, R, hS, S) - what MFGSolutionsFixedTrajectory.simulateAllProcesses records step by step on pre-drawn increments
 * (coupledMFG/MFGSolutions.py:23-57, getAllStates MFGModel.py:106-107). */
FBSDEJ_API int fbsdej_solver_mfg_states(fbsdej_solver* s, int B, float* out);
/* Forward + hand-derived adjoint: trainOpt's tape.gradient (SolversJumpDiff.py:47-53, MFGSolvers.py:50-73).
 * out: 4 + P floats. */
FBSDEJ_API int fbsdej_solver_grad(fbsdej_solver* s, const float* theta, int B, int B_global, float* out);

/* Keras-form Adam (optimizers.Adam, SolversJumpDiff.py:55).  `t_dev` is a device int32 step
 * counter incremented by the call (so the launch sequence is CUDA-graph replayable).  mask may be NULL. */
FBSDEJ_API int fbsdej_adam_step(fbsdej_ctx* ctx, float* theta, float* m, float* v, const float* grad, const float* mask, int n,
                     float lr, float beta1, float beta2, float eps, int* t_dev);

/* Data-parallel building block: simulate (Philox counter word 0 = path_offset + local path id, iteration read from
 * the device counter iter_dev) followed by fbsdej_solver_grad.  The caller all-reduces `out` (4 + P floats, sum)
 * over the ranks, then calls fbsdej_adam_step on out + 4 and fbsdej_bump_u32(iter_dev) on every rank. */
FBSDEJ_API int fbsdej_solver_grad_step(fbsdej_solver* s, const float* theta, uint64_t seed, const uint32_t* iter_dev,
                            uint32_t path_offset, int B, int B_global, float* out);
FBSDEJ_API int fbsdej_bump_u32(fbsdej_ctx* ctx, uint32_t* p);

/* n_steps x (simulate -> grad -> Adam) for single-GPU training, captured once as a CUDA graph and
 * replayed: the inner `for epoch in range(num_epoch): trainOpt(...)` loop (SolversJumpDiff.py:62-64).
 * iter_dev: device uint32 iteration counter (Philox counter word), incremented per step.
 * loss_out (device, may be NULL) receives the n_steps losses. */
FBSDEJ_API int fbsdej_solver_train_steps(fbsdej_solver* s, float* theta, float* m, float* v, const float* mask, int* t_dev,
                              uint32_t* iter_dev, uint64_t seed, int B, int n_steps, float lr, float beta1,
                              float beta2, float eps, float* loss_out);

/* Data-parallel training without a host-driven collective (replaces "grad_step -> NCCL all_reduce -> adam_step" of the
 * reference-side loop; the reference itself is single-device).  One rank per GPU (one process each), or several
 * solvers of one process.
 *   dp_init     allocates this rank's exchange buffer and returns its 64-byte CUDA IPC handle;
 *   dp_buffer   the same buffer as a raw device pointer (for ranks that share a process);
 *   dp_connect  handles: world x 64 bytes gathered from all ranks (may be NULL if raw_ptrs covers every peer);
 *               raw_ptrs: world entries, non-NULL for peers of this process (may be NULL);
 *   train_steps_dp  n_steps training steps on the shard [path_offset, path_offset + B) of a global batch of B_global paths:
 *               the kernel that finishes a step writes its [loss | gradient] vector into every peer's buffer (NVLink peer
 *               stores), waits for the peers' vectors, adds them in rank order and applies the same Adam update on every
 *               rank.  Same CUDA graph replay as train_steps; all ranks must call it with the same n_steps.
 *   dp_check    synchronises the ctx stream; 0, or -3 (message in fbsdej_last_error) if an exchange timed out: a peer did not
 *               arrive within FBSDEJ_DP_TIMEOUT_MS (default 30 s).  The step that timed out - and every later one - is VOID on
 *               this rank (parameters, Adam slots and counters untouched; the recorded loss reads NaN), and the error word is
 *               raised in every reachable rank's buffer, so all ranks fail the check instead of diverging silently. */
FBSDEJ_API int fbsdej_solver_dp_init(fbsdej_solver* s, int rank, int world, unsigned char* handle64);
FBSDEJ_API int fbsdej_solver_dp_buffer(fbsdej_solver* s, void** ptr);
FBSDEJ_API int fbsdej_solver_dp_connect(fbsdej_solver* s, const unsigned char* handles, void* const* raw_ptrs);
FBSDEJ_API int fbsdej_solver_dp_check(fbsdej_solver* s);
FBSDEJ_API int fbsdej_solver_train_steps_dp(fbsdej_solver* s, float* theta, float* m, float* v, const float* mask, int* t_dev,
                                 uint32_t* iter_dev, uint64_t seed, int B, int B_global, uint32_t path_offset, int n_steps,
                                 float lr, float beta1, float beta2, float eps, float* loss_out);

/* Per-kernel device times of one training iteration, measured with CUDA events on the ctx stream around each launch
 * (bench.py's roofline numbers).  Runs `reps` iterations WITHOUT the Adam update (theta unchanged) and writes the mean
 * milliseconds: ms[0] simulate paths, ms[1] simulate + compact compensator samples, ms[2] forward, ms[3] backward,
 * ms[4] partial reduction. */
FBSDEJ_API int fbsdej_solver_profile(fbsdej_solver* s, const float* theta, uint64_t seed, int B, int reps, float* ms_host);

/* Generic row-wise network evaluation: Net.call (Networks.py:17-23).  x [rows][nin] row-major,
 * y [rows][nout] row-major.  net_index selects the net inside theta's flat layout. */
FBSDEJ_API int fbsdej_solver_net_forward(fbsdej_solver* s, const float* theta, int net_index, const float* x, int rows, float* y);
/* Same, without a solver object: theta_net points at ONE net's parameters (W1,b1,W2,b2,W3,b3 of nin->H->H->nout). */
FBSDEJ_API int fbsdej_net_forward(fbsdej_ctx* ctx, const float* theta_net, int nin, int H, int L, int nout, int act, const float* x,
                       int rows, float* y);
/* Closed-form / FFT price A(iStep, X) (pricingModels.py:40-49 / :156-179); X [d][n] component planes. */
FBSDEJ_API int fbsdej_solver_price(fbsdej_solver* s, int iStep, const float* X, int n, float* out);

/* Layout helpers: [N][B][d] (reference-style, d innermost) <-> [N][d][B]. */
FBSDEJ_API int fbsdej_transpose_nbd_to_ndb(fbsdej_ctx* ctx, const float* src, float* dst, int N, int B, int d);
FBSDEJ_API int fbsdej_transpose_ndb_to_nbd(fbsdej_ctx* ctx, const float* src, float* dst, int N, int B, int d);

/* Number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
FBSDEJ_API long long fbsdej_ctx_launch_count(const fbsdej_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif

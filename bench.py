#!/usr/bin/env python
"""Benchmark of the deep-FBSDE-with-jumps training hot path (BASELINE.json metric: train iters/s & path-steps/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # reference-equivalent CPU path (torch restatement)

A "step" = one training iteration of the workload: simulate increments -> forward -> adjoint -> reduce ->
(exchange) -> Adam.  Workloads (config.workload):
  default     SURVEY 8d config 3 per GPU: Merton d=10 geometric basket, N=100 time steps, H=21 tanh nets, SolverGlobalSumLocalReg,
              2^16 paths on EVERY GPU (weak scaling: the global batch is N x 2^16; N = 1 is config 3 itself), so that the
              1 -> 8 GPU curve compares like with like whichever way the N = 1 point is launched
  --config 5  the same model at a fixed global batch of 2^20 paths sharded over the N GPUs (strong scaling; N = 1 fits: 9.6 GB)
  --config 1 | 2 | 4   the reference's own defaults (mainMerton.py, mainVG.py, mainMFGComparison.py) through the drop-in classes
              on one GPU: metric = training iterations/s (path-steps/s beside it); --solver picks the solver class
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import math
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

MERTON = dict(T=1.0, N=100, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)
D, H_WIDTH, LIMIT, ALIN, LR = 10, 21, 100, 0.1, 3e-4
SOLVERS = {"SumLocalReg": 0, "MultiStepReg": 0, "Global": 256}     # name -> default compensator samples M


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="SURVEY 8d configuration (default 3: 2^16 paths per GPU; 5: 2^20 paths in total; 1 / 2 / 4: reference defaults)")
    ap.add_argument("--paths", type=int, default=0, help="global Monte-Carlo batch (default 2^16 per GPU; --config 5: 2^20)")
    ap.add_argument("--solver", default="", help="configs 3 / 5: SumLocalReg (default), MultiStepReg, Global; configs 1 / 2 / 4: a "
                                                 "solver class name of the reference (default SolverGlobalFBSDE)")
    ap.add_argument("--M", type=int, default=-1, help="compensator samples for --solver Global (default 256)")
    ap.add_argument("--cpu-paths", type=int, default=0,
                    help="paths of the bounded CPU sample per step (0 = automatic: 16384 for the cpu_baseline leg; for --impl "
                         "reference the largest power of two <= the batch that keeps the run near two minutes - the CPU "
                         "restatement is more efficient on larger samples)")
    ap.add_argument("--mma", default="tcgen05", choices=["ffma", "tcgen05"], help="layer arithmetic of the fused kernels")
    ap.add_argument("--hidden", type=int, default=21, help="hidden width of the networks (configs 3 / 5; 21 = the reference's; > 22 "
                                                           "runs the fp32 FFMA kernels: SURVEY 8d's H = 32 variant)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def global_batch(a, world):
    """Default: 2^16 paths per GPU (config 3 on every GPU, weak scaling); --config 5: 2^20 paths in total (strong scaling)."""
    if a.paths:
        return a.paths
    return 2 ** 20 if a.config == 5 else 2 ** 16 * world


def workload_config(a, B, world):
    M = SOLVERS[a.solver] if a.M < 0 else a.M
    which = ("config 5: fixed global batch, sharded" if a.config == 5 else
             "config 3" if world == 1 else "config 3 on every GPU: %d x %d paths" % (world, B // world))
    name = ("Merton d=10 geometric basket, N=100, H=%d tanh, Solver%s%s, B=%d paths (SURVEY 8d %s)"
            % (H_WIDTH, "Global" + a.solver if a.solver.endswith("Reg") else a.solver + "FBSDE", "" if M == 0 else " M=%d" % M, B, which))
    return M, {"workload": name, "paths": B, "paths_per_gpu": B // world, "time_steps": MERTON["N"], "d": D, "hidden": H_WIDTH,
               "solver": a.solver,
               "compensator_M": M, "mma": a.mma, "parallelism": "dp%d" % world,
               "l2_policy": "working set exceeds L2: %.0f MB of per-path-step records per rank are written by the forward sweep and "
                            "read back by the adjoint every step (126 MB L2); no flush needed"
               % ((2 * D + 3) * MERTON["N"] * (B // world) * 4 / 1e6)}


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_iteration_factory(a, M, B_cpu):
    """One training iteration of the reference-equivalent CPU restatement (oracle/, torch eager + autograd + Keras-form
    Adam), noise drawn on the CPU the way the reference draws it."""
    import torch
    import helpers as H
    from oracle import MertonOracle, KerasAdam, pricing_loss
    from oracle.pricing import sample_pricing_noise
    om = MertonOracle(aLin=ALIN, limit=LIMIT, d=D, **MERTON)
    layout = H.pricing_layout("merton", a.solver, D, H_WIDTH)
    theta = torch.tensor(H.random_theta(layout, 0), requires_grad=True)
    opt = KerasAdam(layout.total, LR)
    gen = torch.Generator().manual_seed(0)

    def it():
        noise = sample_pricing_noise(om, a.solver, B_cpu, max(M, 1), gen)
        if theta.grad is not None:
            theta.grad = None
        loss = pricing_loss(om, a.solver, layout, theta, noise, B_cpu)
        loss.backward()
        opt.step(theta.data, theta.grad)
        return float(loss.detach())
    return it


def run_reference(a):
    """--impl reference: TensorFlow (the reference's runtime) is not installable here, so the reference arm is the
    CPU restatement of the same solver on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every host core whichever way it is launched
    torch.set_num_threads(os.cpu_count() or 1)
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    B = global_batch(a, world)
    M, cfg = workload_config(a, B, world)
    # the same bounded sample at every N (the restatement's throughput depends on the sample size, not on the batch it stands for)
    cpu_paths = a.cpu_paths if a.cpu_paths > 0 else 16384
    it = cpu_iteration_factory(a, M, cpu_paths)
    for _ in range(max(1, min(a.warmup, 1))):
        it()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        it()
    dt = (time.perf_counter() - t0) / a.steps
    val = cpu_paths * MERTON["N"] / dt
    sample = "%d-path sample of the %d-path batch per step, all %d time steps, fwd+bwd+Adam" % (cpu_paths, B, MERTON["N"])
    print(json.dumps({
        "impl": "reference", "metric": "path-steps/s", "value": val, "unit": "path-steps/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "iters_per_s_equiv": val / (B * MERTON["N"]), "higher_is_better": True,
        "scaling": "strong" if a.config == 5 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "path-steps/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference-equivalent CPU restatement (torch eager), not TensorFlow"}))


# ---------------------------------------------------------------------------------------------------------------
def run_native(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from deepfbsdejsolvers_b200 import Context, set_seed
    from deepfbsdejsolvers_b200 import coupledPricing as cp
    from deepfbsdejsolvers_b200.solver_base import shard, attach_peers
    import deepfbsdejsolvers_b200._lib as L

    B = global_batch(a, world)
    M, cfg = workload_config(a, B, world)
    off, Bl = shard(B, rank, world)
    ctx = Context.default(local)
    set_seed(0)
    mm = cp.MertonJumpModel(MERTON["T"], MERTON["N"], MERTON["r"], MERTON["muJ"], MERTON["sigmaJ"], MERTON["sigma"], MERTON["lam"],
                            MERTON["K"], MERTON["x0"], cp.AbsCoupling(ALIN), LIMIT, d=D)
    layer = [H_WIDTH, H_WIDTH]
    if a.solver == "Global":
        solver = cp.SolverGlobalFBSDE(mm, cp.Net(1, D, layer, "tanh"), cp.Net(0, 1, layer, "tanh"), LR, M=M)
    elif a.solver == "SumLocalReg":
        solver = cp.SolverGlobalSumLocalReg(mm, cp.Net(0, 1, layer, "tanh"), cp.Net(0, 1, layer, "tanh"), LR,
                                            tensor_cores=a.mma == "tcgen05")
    else:
        solver = cp.SolverGlobalMultiStepReg(mm, cp.Net(0, 1, layer, "tanh"), cp.Net(0, 1, layer, "tanh"), LR,
                                             tensor_cores=a.mma == "tcgen05")
    s = solver.build()
    N = MERTON["N"]

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        ctx.sync()

    # N > 1: the step's [loss | gradient] exchange runs inside the finishing kernel over NVLink peer memory
    # (fbsdej_solver_train_steps_dp: one CUDA graph per step, no collective call); FBSDEJ_DP=nccl selects the
    # grad_step -> NCCL all_reduce -> adam_step loop instead
    p2p = world > 1 and os.environ.get("FBSDEJ_DP", "p2p") != "nccl"
    if p2p:
        p2p = attach_peers(s, dist, rank, world)
    if world > 1:
        cfg["dp_exchange"] = ("inside the finishing kernel over NVLink peer memory (CUDA IPC buffers), no collective call" if p2p
                              else "NCCL all_reduce of [loss | gradient] per step")

    def step():
        if world == 1:
            s.train_steps(0, B, 1, LR)
        elif p2p:
            s.train_steps_dp(0, Bl, B, off, 1, LR)
        else:
            out = s.grad_step(0, Bl, B, off)
            with torch.cuda.stream(ctx.stream):
                dist.all_reduce(out)
            s.adam_step(LR)
            s.bump_iteration()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(ctx.stream)
        for _ in range(n):
            fn()
        e1.record(ctx.stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=ctx.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launches
    ms = timed(step, a.steps)
    launches = ctx.launches - l0
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms / a.steps
    value = B * N / (ms_step * 1e-3)

    # ---- end to end through the C-ABI with HOST buffers: H2D of the step's increments + step + D2H of the loss ----
    # The step's Brownian / jump increments live in pinned host memory; a copy stream uploads step k+1 into the other
    # device buffer while step k computes (both inside the timed region), the loss comes back to the host every step.
    e2e, e2e_rng = None, None
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    if not a.no_e2e and world == 1:
        nfl = N * D * Bl
        # one step's increments in pinned host memory (synthetic; the same block is uploaded for every step): Brownian planes
        # dense, compound-Poisson jump planes as their non-zero entries - exactly 0 wherever no jump fell into the step
        dt = MERTON["T"] / N
        gen = torch.Generator().manual_seed(1234 + rank)
        dW_h = torch.randn(nfl, dtype=torch.float32, generator=gen).mul_(dt ** 0.5).pin_memory()
        hit = torch.rand(nfl, generator=gen) < (1.0 - math.exp(-MERTON["lam"] * dt))
        jidx_h = torch.nonzero(hit).flatten().to(torch.int32).pin_memory()      # < 2^31 entries per rank
        nnz = int(jidx_h.numel())
        jval_h = (torch.randn(nnz, dtype=torch.float32, generator=gen) * MERTON["sigmaJ"] + MERTON["muJ"]).pin_memory()
        del hit
        dev = [[ctx.empty(nfl), ctx.empty(max(nnz, 1), dtype=torch.int32), ctx.empty(max(nnz, 1))] for _ in range(2)]
        jmc = ctx.zeros(N * D * max(M, 1)) if M > 0 else None
        copy_stream = torch.cuda.Stream(ctx.device)
        up = [torch.cuda.Event(), torch.cuda.Event()]        # upload of buffer b finished
        free = [torch.cuda.Event(), torch.cuda.Event()]      # compute on buffer b finished
        k = [0]

        def upload(b):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])
                dev[b][0].copy_(dW_h, non_blocking=True)
                dev[b][1][:nnz].copy_(jidx_h, non_blocking=True)
                dev[b][2][:nnz].copy_(jval_h, non_blocking=True)
                up[b].record(copy_stream)

        def e2e_step():
            b = k[0] & 1
            k[0] += 1
            upload(b ^ 1)                                    # next step's increments -> the other buffer
            ctx.stream.wait_event(up[b])
            L.check(L.lib.fbsdej_solver_set_noise_sparse_jumps(s.handle, Bl, dev[b][0].data_ptr(), dev[b][1].data_ptr(),
                                                               dev[b][2].data_ptr(), nnz,
                                                               jmc.data_ptr() if jmc is not None else None))
            L.check(L.lib.fbsdej_solver_grad(s.handle, s.theta.data_ptr(), Bl, B, s.out.data_ptr()))
            free[b].record(ctx.stream)
            with torch.cuda.stream(ctx.stream):
                if world > 1:
                    dist.all_reduce(s.out)
            s.adam_step(LR)
            with torch.cuda.stream(ctx.stream):
                loss_host.copy_(s.out[:1], non_blocking=True)
            ctx.sync()

        for b in range(2):
            free[b].record(ctx.stream)
        upload(0)
        for _ in range(2):
            e2e_step()
        ms_e = timed(e2e_step, a.steps) / a.steps
        copy_stream.synchronize()
        e2e = {"value": B * N / (ms_e * 1e-3), "unit": "path-steps/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": (nfl * 4 + nnz * 8) * world, "d2h_bytes_per_step": 4 * world,
               "what": "host (pinned) increments of the step - Brownian planes dense, compound-Poisson jump planes as (index, "
                       "value) of their non-zero entries - -> H2D (double-buffered on a copy stream) -> "
                       "fbsdej_solver_set_noise_sparse_jumps -> fbsdej_solver_grad -> fbsdej_adam_step -> loss D2H + sync "
                       "every step; bounded by the PCIe upload of 4*d bytes per path-step"}
        del dev

    if not a.no_e2e:
        # the production call: Solver.train_steps draws the increments on the device (Philox), so its per-step host input is
        # (seed, step count, learning rate) and its per-step host output the loss
        loss_dev = ctx.zeros(1)

        def rng_step():
            s.train_steps(0, B, 1, LR, loss_out=loss_dev)
            with torch.cuda.stream(ctx.stream):
                loss_host.copy_(loss_dev, non_blocking=True)
            ctx.sync()

        def rng_step_dp():
            if p2p:
                s.train_steps_dp(0, Bl, B, off, 1, LR, loss_out=loss_dev)
            else:
                out = s.grad_step(0, Bl, B, off)
                with torch.cuda.stream(ctx.stream):
                    dist.all_reduce(out)
                    loss_dev.copy_(out[:1])
                s.adam_step(LR)
                s.bump_iteration()
            with torch.cuda.stream(ctx.stream):
                loss_host.copy_(loss_dev, non_blocking=True)
            ctx.sync()

        one = rng_step if world == 1 else rng_step_dp
        for _ in range(2):
            one()
        ms_r = timed(one, a.steps) / a.steps
        e2e_rng = {"value": B * N / (ms_r * 1e-3), "unit": "path-steps/s", "ms_per_step": ms_r, "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 4 * world,
                   "what": ("fbsdej_solver_train_steps(n_steps=1)" if world == 1 else "fbsdej_solver_train_steps_dp(n_steps=1) on every rank")
                   + " + loss D2H + host sync every step; increments drawn on the device, as the reference draws them inside its "
                     "graph: the per-step host input of the product call is (seed, learning rate)"}

    # ---- per-kernel device times + roofline (rank 0) ----------------------------------------------------------------
    roof, kernels = None, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        prof = s.profile(0, Bl, reps=max(3, min(a.steps, 10)))
        ps = Bl * N
        jump = M > 0
        # the tcgen05 Merton solvers draw the increments inside the forward sweep: no simulation kernel, and the fused
        # kernel is charged with the simulation's algorithmic bytes as well ("fused kernel" paragraph)
        fused_sim = a.mma == "tcgen05" and not jump and os.environ.get("FBSDEJ_NO_FUSED_RNG") is None
        # algorithmic bytes per path-step (SURVEY 8d / BASELINE.md section 5); Reg solvers have no Z / Gam / comp words
        alg = {"sim_paths": 8 * D * ps,
               "forward": 4 * ((5 * D + 4) if jump else (3 * D + 2 + D + 1)) * ps,
               "backward": 4 * ((7 * D + 4) if jump else (5 * D + 2 + 2 * D + 2 - D)) * ps}
        if fused_sim:
            alg["forward"] += alg.pop("sim_paths")
        kernels = {}
        if fused_sim:
            kernels["sim_paths"] = {"ms": 0.0, "fused_into": "forward (reg_forward_tc draws the increments in registers)"}
        for kname, b in alg.items():
            t = prof[kname]
            kernels[kname] = {"ms": t, "algorithmic_bytes": b, "achieved_GBps": b / (t * 1e-3) / 1e9 if t > 0 else None,
                              "frac_of_measured_hbm": b / (t * 1e-3) / 1e9 / peak if t > 0 else None}
        kernels["reduce"] = {"ms": prof["reduce"]}
        if jump:
            kernels["sim_compensator"] = {"ms": prof["sim_compensator"]}
        dom = max(alg, key=lambda n: prof[n])
        tc = a.mma == "tcgen05" and not jump
        kname = {"sim_paths": "sim_merton_kernel", "forward": "reg_forward_tc" if tc else "pricing_forward",
                 "backward": "reg_backward_tc" if tc else "pricing_backward"}[dom]
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (config 3)
            prof_js = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if Bl == prof_js["paths"] and N == prof_js["time_steps"]:
                traffic = prof_js["kernels"][kname]["dram_bytes"]
        except Exception:
            pass
        roof = {"kernel": kname,
                "bound": "hbm", "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_measured_hbm"], "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "share_of_step": prof[dom] / sum(prof.values()),
                "algorithmic_bytes": kernels[dom]["algorithmic_bytes"],
                "note": "fused kernel: the network GEMMs (tcgen05), tanh, the closed-form coupling and the adjoint run inside "
                        "the launch, so it is bound by the tensor / MUFU pipes and the per-step MMA round trips, not by HBM; "
                        "frac = algorithmic path-tensor bytes (SURVEY 8d) / launch time / measured copy bandwidth"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_paths = a.cpu_paths if a.cpu_paths > 0 else min(B, 16384)
        it = cpu_iteration_factory(a, M, cpu_paths)
        it()
        t0, n = time.perf_counter(), 0
        while n < 2 or (time.perf_counter() - t0 < 15.0 and n < 50):
            it(); n += 1
        dt = (time.perf_counter() - t0) / n
        cpu = {"value": cpu_paths * N / dt, "unit": "path-steps/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d iterations of a %d-path sample of the batch, all %d time steps, fwd+bwd+Adam (torch eager restatement "
                         "of the reference solver, not TensorFlow)" % (n, cpu_paths, N), "ms_per_iteration": dt * 1e3}

    in_sync = None
    if world > 1:        # every rank applied the same updates: the parameter vectors must agree bit for bit
        with torch.cuda.stream(ctx.stream):
            chk = torch.stack([s.theta.double().sum(), s.theta.double().abs().sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ctx.sync()
        in_sync = bool(torch.equal(lo, hi))
        cfg["ranks_in_sync"] = in_sync
    if rank == 0:
        line = {"metric": "path-steps/s", "value": value, "unit": "path-steps/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "iters_per_s": 1e3 / ms_step, "higher_is_better": True,
                "scaling": "strong" if a.config == 5 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                # N > 1: the end-to-end number is the product call (device-drawn increments); the host-increment pipeline of the
                # N = 1 line would measure N ranks sharing the host's PCIe, not the library
                "config": cfg, "clocks": clk, "e2e": e2e if e2e is not None else e2e_rng, "e2e_device_rng": e2e_rng, "gpu_launches": int(launches), "roofline": roof,
                "kernels": kernels, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    H_WIDTH = args.hidden
    if H_WIDTH > 22:
        args.mma = "ffma"
    if args.config in (1, 2, 4):
        from bench_reference_defaults import run_config       # scripts/: configs 1 / 2 / 4
        run_config(args)
        sys.exit(0)
    args.solver = args.solver or "SumLocalReg"
    if args.solver not in SOLVERS:
        sys.exit("--solver must be one of %s for configs 3 / 5" % list(SOLVERS))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)

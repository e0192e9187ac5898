"""bench.py --config 1 | 2 | 4: BASELINE.json configs[0], [1], [3] - the reference's own defaults (mainMerton.py, mainVG.py,
mainMFGComparison.py: batch sizes, 5000 compensator samples, network widths, learning rates) through the drop-in solver classes
on one B200.  One JSON line in bench.py's format:

  metric / value   training iterations/s, device-timed (CUDA events around K graph-replayed steps); path_steps_per_s beside it
  e2e              the call a user of the reference makes: Solver.train(batchSize, batchSizeVal, num_epoch = K, num_epochExt = 1) -
                   wall clock per iteration, validation pass and Y0 report included; its host inputs are (seed, learning rate)
  cpu_baseline     the reference-equivalent CPU restatement (oracle/, torch eager + autograd + Keras-form Adam) on the same shapes,
                   noise drawn on the CPU the way the reference draws it
  roofline         these shapes (10 or 128 paths) fill a handful of CTAs; the step is a chain of MMA round trips, so the bound is
                   LATENCY: floor = N time steps x (2 forward + 4 adjoint) layer round trips x the measured round-trip time of a
                   lone CTA (scripts/mma_microbench.cu -> profiles/r2_mma_microbench.csv), frac = floor / measured
"""
from __future__ import annotations

import contextlib
import csv
import io
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCHEME_OF = {"SolverGlobalFBSDE": "Global", "SolverMultiStepFBSDE1": "MultiStep1", "SolverMultiStepFBSDE2": "MultiStep2",
             "SolverSumLocalFBSDE1": "SumLocal1", "SolverSumLocalFBSDE2": "SumLocal2", "SolverGlobalSumLocalReg": "SumLocalReg",
             "SolverGlobalMultiStepReg": "MultiStepReg", "SolverMultiStepFBSDE": "MultiStep", "SolverSumLocalFBSDE": "SumLocal"}


def make_case(config: int, name: str, ctx):
    """(solver object, train batchSize, batchSizeVal, effective train batch, N, M, description) of one reference default."""
    import helpers as H
    from deepfbsdejsolvers_b200 import coupledPricing as cp, coupledMFG as cm
    from deepfbsdejsolvers_b200.coupledPricing import SolversPureJump as pj

    def net(bY0, nout):
        return cp.Net(bY0, nout, [21, 21], "tanh")
    reg = name.endswith("Reg")
    if config == 1:       # mainMerton.py:13-25, 57, 94-118
        M = H.MERTON
        mm = cp.MertonJumpModel(M["T"], M["N"], M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(0.1), 30)
        mk = {"SolverGlobalFBSDE": lambda: cp.SolverGlobalFBSDE(mm, net(1, 1), net(0, 1), 4e-4, ctx=ctx),
              "SolverMultiStepFBSDE1": lambda: cp.SolverMultiStepFBSDE1(mm, net(0, 2), 3e-4, ctx=ctx),
              "SolverMultiStepFBSDE2": lambda: cp.SolverMultiStepFBSDE2(mm, net(0, 2), net(0, 1), 3e-4, ctx=ctx),
              "SolverSumLocalFBSDE1": lambda: cp.SolverSumLocalFBSDE1(mm, net(0, 2), 3e-4, ctx=ctx),
              "SolverSumLocalFBSDE2": lambda: cp.SolverSumLocalFBSDE2(mm, net(0, 2), net(0, 1), 3e-4, ctx=ctx),
              "SolverGlobalSumLocalReg": lambda: cp.SolverGlobalSumLocalReg(mm, net(0, 1), net(0, 1), 3e-4, ctx=ctx),
              "SolverGlobalMultiStepReg": lambda: cp.SolverGlobalMultiStepReg(mm, net(0, 1), net(0, 1), 3e-4, ctx=ctx)}[name]
        return mk(), 10, 100, 10000 if reg else 10, M["N"], 0 if reg else 5000, "mainMerton.py defaults: Merton 1D, N=50, H=21 tanh"
    if config == 2:       # mainVG.py:12-24, 54, 88-111
        V = H.VG
        mm = cp.VGmodel(V["T"], V["N"], V["r"], V["theta"], V["kappa"], V["sigmaJ"], V["K"], V["x0"], cp.AbsCoupling(0.1))
        mk = {"SolverGlobalFBSDE": lambda: pj.SolverGlobalFBSDE(mm, net(0, 1), net(1, 1), 5e-4, ctx=ctx),
              "SolverMultiStepFBSDE1": lambda: pj.SolverMultiStepFBSDE1(mm, net(0, 1), 3e-4, ctx=ctx),
              "SolverMultiStepFBSDE2": lambda: pj.SolverMultiStepFBSDE2(mm, net(0, 1), net(0, 1), 3e-4, ctx=ctx),
              "SolverSumLocalFBSDE1": lambda: pj.SolverSumLocalFBSDE1(mm, net(0, 1), 3e-4, ctx=ctx),
              "SolverSumLocalFBSDE2": lambda: pj.SolverSumLocalFBSDE2(mm, net(0, 1), net(0, 1), 3e-4, ctx=ctx),
              "SolverGlobalSumLocalReg": lambda: pj.SolverGlobalSumLocalReg(mm, net(0, 1), net(0, 1), 1.5e-4, ctx=ctx),
              "SolverGlobalMultiStepReg": lambda: pj.SolverGlobalMultiStepReg(mm, net(0, 1), net(0, 1), 1.5e-4, ctx=ctx)}[name]
        return mk(), 10, 100, 10000 if reg else 10, V["N"], 0 if reg else 5000, "mainVG.py defaults: Variance Gamma 1D, N=30, H=21 tanh"
    # mainMFGComparison.py:13-33, 92-94, 108-138 (nbDays = 2: N = 95; couplage ON)
    P = H.mfg_params()
    mm = cm.ModelCoupledFBSDE(**P)
    wh, wi = {"SolverGlobalFBSDE": (2, 3), "SolverMultiStepFBSDE": (3, 4), "SolverSumLocalFBSDE": (3, 4),
              "SolverGlobalSumLocalReg": (1, 1), "SolverGlobalMultiStepReg": (1, 1)}[name]
    method = {"SolverGlobalFBSDE": "Global", "SolverMultiStepFBSDE": "SumMultiStep", "SolverSumLocalFBSDE": "SumLocal",
              "SolverGlobalSumLocalReg": "SumLocalReg", "SolverGlobalMultiStepReg": "SumMultiStepReg"}[name]
    km = cm.kerasModels(cm.Net_hat, cm.Net, method, wh, wi, [20, 20], [22, 22], "tanh", "tanh")
    lr = 1e-3 if name == "SolverGlobalFBSDE" else (1e-4 if reg else 1.5e-4)
    return (getattr(cm, name)(mm, km, lr, "ON", ctx=ctx), 128, 1280, 128, mm.N, 0,
            "mainMFGComparison.py defaults: smart-grid MFG, Cox jumps, N=95, nets 4-20-20 / 6-22-22 tanh, couplage ON")


def cpu_iteration(config: int, name: str, B: int, M: int):
    import torch
    import helpers as H
    from oracle import MertonOracle, VGOracle, MFGOracle, KerasAdam, pricing_loss, mfg_loss
    from oracle.pricing import sample_pricing_noise
    from oracle.mfg import sample_mfg_noise
    gen = torch.Generator().manual_seed(0)
    scheme = SCHEME_OF[name]
    if config == 4:
        model = MFGOracle(**H.mfg_params())
        layout = H.mfg_layout(scheme)
    else:
        model = MertonOracle(aLin=0.1, limit=30, d=1, **H.MERTON) if config == 1 else VGOracle(aLin=0.1, **H.VG)
        layout = H.pricing_layout("merton" if config == 1 else "vg", scheme, 1)
    theta = torch.tensor(H.random_theta(layout, 0), requires_grad=True)
    opt = KerasAdam(layout.total, 3e-4)

    def it():
        theta.grad = None
        if config == 4:
            lh, li = mfg_loss(model, scheme, layout, theta, sample_mfg_noise(model, B, gen), B)
            loss = lh + li
        else:
            loss = pricing_loss(model, scheme, layout, theta, sample_pricing_noise(model, scheme, B, max(M, 1), gen), B)
        loss.backward()
        opt.step(theta.data, theta.grad)
    return it


def roundtrip_us(sm_mhz: float):
    """Measured round trip of one layer evaluation on a lone CTA (profiles/r2_mma_microbench.csv, 6-MMA row, one CTA)."""
    try:
        for r in csv.reader(open(os.path.join(ROOT, "profiles", "r2_mma_microbench.csv"))):
            if r and r[0].startswith("round trip") and "+ 6 tf32" in r[0] and r[4] == "1":
                return float(r[5]) / sm_mhz
    except Exception:
        pass
    return None


def run_config(a):
    import torch
    from bench import ClockSampler
    from deepfbsdejsolvers_b200 import Context, set_seed
    torch.cuda.set_device(0)
    ctx = Context.default(0)
    set_seed(0)
    name = a.solver or "SolverGlobalFBSDE"
    solver, bs, bsv, B, N, M, what = make_case(a.config, name, ctx)
    if a.impl == "reference":
        torch.set_num_threads(os.cpu_count() or 1)
        it = cpu_iteration(a.config, name, B, M)
        it()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            it()
        dt = (time.perf_counter() - t0) / a.steps
        print(json.dumps({"impl": "reference", "metric": "train iters/s", "value": 1.0 / dt, "unit": "iters/s", "n_gpus": 1, "steps": a.steps,
                          "warmup": 1, "ms_per_step": dt * 1e3, "path_steps_per_s": B * N / dt, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "%s, %s, B=%d, M=%d (SURVEY 8d config %d)" % (what, name, B, M, a.config)},
                          "cpu_baseline": {"value": 1.0 / dt, "unit": "iters/s", "cores": torch.get_num_threads(), "kind": "port",
                                           "sample": "full training iterations of the reference shapes"},
                          "e2e": {"value": 1.0 / dt, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    s = solver.build()
    lr = solver.lRate
    warm = max(a.warmup, 3)
    s.reset_optimizer()
    s.train_steps(0, B, warm + 2, lr)          # (+ graph capture)
    ctx.sync()
    clocks = ClockSampler(0)
    clocks.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    s.train_steps(0, B, a.steps, lr)
    e1.record(ctx.stream)
    ctx.sync()
    launches = ctx.launches - l0
    clk = clocks.stop()
    ms = e0.elapsed_time(e1) / a.steps
    # ---- end to end: the reference's own call -------------------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        with contextlib.redirect_stdout(io.StringIO()):
            solver.train(bs, bsv, 5, 1)
            t0 = time.perf_counter()
            solver.train(bs, bsv, a.steps, 1)
            wall = time.perf_counter() - t0
        e2e = {"value": a.steps / wall, "unit": "iters/s", "ms_per_step": wall * 1e3 / a.steps, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 8.0 / a.steps,
               "what": "%s.train(%d, %d, num_epoch=%d, num_epochExt=1): wall clock per training iteration incl. the validation pass "
                       "(batch %d) and the Y0 report of the outer epoch; increments are drawn on the device, host input = (seed, lr)"
                       % (name, bs, bsv, a.steps, bsv)}
    prof = s.profile(0, B, reps=10)
    sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
    rt = roundtrip_us(sm_mhz)
    roof = None
    if rt is not None:
        floor_ms = N * 6 * rt * 1e-3
        roof = {"kernel": "forward + adjoint sweeps (one chain of N steps x 6 layer round trips)", "bound": "latency",
                "achieved": 1e3 / ms, "peak": 1e3 / floor_ms, "unit": "iters/s", "frac": floor_ms / ms, "traffic": None,
                "round_trip_us": rt, "floor_ms": floor_ms,
                "note": "%d paths%s occupy a handful of CTAs (147 of 148 SMs idle for the MFG solvers): neither HBM nor the tensor pipe "
                        "can bound the step; the floor is the serial chain of tensor-memory store -> barrier -> MMAs -> commit -> "
                        "mbarrier -> tensor-memory load round trips, measured on a lone CTA" % (B, " x %d compensator samples" % M if M else "")}
    cpu = None
    if not a.no_cpu_baseline:
        it = cpu_iteration(a.config, name, B, M)
        it()
        t0, n = time.perf_counter(), 0
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
            it(); n += 1
        dt = (time.perf_counter() - t0) / n
        cpu = {"value": 1.0 / dt, "unit": "iters/s", "cores": torch.get_num_threads(), "kind": "port", "ms_per_iteration": dt * 1e3,
               "sample": "%d full training iterations at the reference shapes (torch eager restatement of the reference solver, "
                         "not TensorFlow)" % n}
    print(json.dumps({
        "metric": "train iters/s", "value": 1e3 / ms, "unit": "iters/s", "n_gpus": 1, "steps": a.steps, "warmup": warm, "ms_per_step": ms,
        "path_steps_per_s": B * N * 1e3 / ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s, %s, B=%d paths%s (SURVEY 8d config %d)" % (what, name, B, ", M=%d" % M if M else "", a.config),
                   "paths": B, "time_steps": N, "compensator_M": M, "solver": name, "mma": "tcgen05" if s_uses_tc(solver) else "ffma",
                   "l2_policy": "latency-bound shapes (working set < 1 MB); nothing to flush"},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
        "kernels": {k: {"ms": v} for k, v in prof.items()}, "cpu_baseline": cpu}))


def s_uses_tc(solver) -> bool:
    try:
        return bool(solver.native.desc_mma_mode)
    except Exception:
        return solver.tensor_cores is not False

#!/usr/bin/env python
"""Training iterations/s of BASELINE.json configs 1, 2 and 4 (the reference's own CPU-runnable defaults) through the drop-in
solver classes on one B200: every solver class of mainMerton.py / mainVG.py / mainMFGComparison.py at the reference's batch
sizes, compensator samples (5000) and network widths.  One JSON line per (config, solver); `--cpu` adds the torch-CPU
restatement (oracle) timed on the host cores for the same shapes.  Parity cases, not bench.py lines (those are config 3/5)."""
import argparse, json, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def time_solver(solver, B, iters, ctx):
    import torch
    s = solver.build()
    from deepfbsdejsolvers_b200.solver_base import TrainLoop
    loop = TrainLoop(s, solver.lRate, 0)
    loop.steps(B, 20)                      # warm-up + graph capture
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    e0.record(ctx.stream)
    s.train_steps(0, B, iters, solver.lRate)
    e1.record(ctx.stream)
    ctx.sync()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated config numbers (default: 1,2,4)")
    a = ap.parse_args()
    import helpers as H
    from deepfbsdejsolvers_b200 import Context, set_seed
    from deepfbsdejsolvers_b200 import coupledPricing as cp, coupledMFG as cm
    ctx = Context.default(0)
    set_seed(0)
    out = []

    def net(bY0, nout):
        return cp.Net(bY0, nout, [21, 21], "tanh")

    # ---- config 1: mainMerton.py defaults (N = 50, B = 10, Reg solvers 10^4, M = 5000) -------------------------------
    M = H.MERTON
    def merton():
        return cp.MertonJumpModel(M["T"], M["N"], M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(0.1), 30)
    cases1 = [("SolverGlobalFBSDE", lambda m: cp.SolverGlobalFBSDE(m, net(1, 1), net(0, 1), 4e-4), 10),
              ("SolverMultiStepFBSDE1", lambda m: cp.SolverMultiStepFBSDE1(m, net(0, 2), 3e-4), 10),
              ("SolverMultiStepFBSDE2", lambda m: cp.SolverMultiStepFBSDE2(m, net(0, 2), net(0, 1), 3e-4), 10),
              ("SolverSumLocalFBSDE1", lambda m: cp.SolverSumLocalFBSDE1(m, net(0, 2), 3e-4), 10),
              ("SolverSumLocalFBSDE2", lambda m: cp.SolverSumLocalFBSDE2(m, net(0, 2), net(0, 1), 3e-4), 10),
              ("SolverGlobalSumLocalReg", lambda m: cp.SolverGlobalSumLocalReg(m, net(0, 1), net(0, 1), 3e-4), 10000),
              ("SolverGlobalMultiStepReg", lambda m: cp.SolverGlobalMultiStepReg(m, net(0, 1), net(0, 1), 3e-4), 10000)]
    only = set(a.only.split(",")) if a.only else {"1", "2", "4"}
    for name, mk, B in (cases1 if "1" in only else []):
        ms = time_solver(mk(merton()), B, a.iters, ctx)
        out.append({"config": "1 mainMerton.py", "solver": name, "paths": B, "time_steps": M["N"], "M": 0 if "Reg" in name else 5000,
                    "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "path_steps_per_s": B * M["N"] * 1e3 / ms})
    # ---- config 2: mainVG.py defaults (N = 30) ---------------------------------------------------------------------------
    V = H.VG
    def vg():
        return cp.VGmodel(V["T"], V["N"], V["r"], V["theta"], V["kappa"], V["sigmaJ"], V["K"], V["x0"], cp.AbsCoupling(0.1))
    from deepfbsdejsolvers_b200.coupledPricing import SolversPureJump as pj
    cases2 = [("SolverGlobalFBSDE", lambda m: pj.SolverGlobalFBSDE(m, net(0, 1), net(1, 1), 5e-4), 10),
              ("SolverMultiStepFBSDE1", lambda m: pj.SolverMultiStepFBSDE1(m, net(0, 1), 3e-4), 10),
              ("SolverMultiStepFBSDE2", lambda m: pj.SolverMultiStepFBSDE2(m, net(0, 1), net(0, 1), 3e-4), 10),
              ("SolverSumLocalFBSDE1", lambda m: pj.SolverSumLocalFBSDE1(m, net(0, 1), 3e-4), 10),
              ("SolverSumLocalFBSDE2", lambda m: pj.SolverSumLocalFBSDE2(m, net(0, 1), net(0, 1), 3e-4), 10),
              ("SolverGlobalSumLocalReg", lambda m: pj.SolverGlobalSumLocalReg(m, net(0, 1), net(0, 1), 1.5e-4), 10000),
              ("SolverGlobalMultiStepReg", lambda m: pj.SolverGlobalMultiStepReg(m, net(0, 1), net(0, 1), 1.5e-4), 10000)]
    for name, mk, B in (cases2 if "2" in only else []):
        ms = time_solver(mk(vg()), B, a.iters, ctx)
        out.append({"config": "2 mainVG.py", "solver": name, "paths": B, "time_steps": V["N"], "M": 0 if "Reg" in name else 5000,
                    "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "path_steps_per_s": B * V["N"] * 1e3 / ms})
    # ---- config 4: mainMFGComparison.py defaults (N = 95, B = 128, couplage ON) -------------------------------------------
    P = H.mfg_params()
    widths = {"SolverGlobalFBSDE": (2, 3), "SolverMultiStepFBSDE": (3, 4), "SolverSumLocalFBSDE": (3, 4),
              "SolverGlobalSumLocalReg": (1, 1), "SolverGlobalMultiStepReg": (1, 1)}
    for name, (wh, wi) in (widths.items() if "4" in only else []):
        mm = cm.ModelCoupledFBSDE(**P)
        method = {"SolverGlobalFBSDE": "Global", "SolverMultiStepFBSDE": "SumMultiStep", "SolverSumLocalFBSDE": "SumLocal",
                  "SolverGlobalSumLocalReg": "SumLocalReg", "SolverGlobalMultiStepReg": "SumMultiStepReg"}[name]
        km = cm.kerasModels(cm.Net_hat, cm.Net, method, wh, wi, [20, 20], [22, 22], "tanh", "tanh")
        solver = getattr(cm, name)(mm, km, 1e-3 if name == "SolverGlobalFBSDE" else 1.5e-4, "ON", ctx=ctx)
        B = 128 * (1 if not name.endswith("Reg") else 1)
        ms = time_solver(solver, B, a.iters, ctx)
        out.append({"config": "4 mainMFGComparison.py", "solver": name, "paths": B, "time_steps": mm.N, "M": 0,
                    "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "path_steps_per_s": B * mm.N * 1e3 / ms})
    if a.cpu:
        # the reference-equivalent CPU restatement (oracle/: torch eager + autograd + Keras-form Adam) on the same shapes, noise
        # drawn on the CPU the way the reference draws it; median of 3 iterations after one warm-up
        import torch
        from oracle import MertonOracle, VGOracle, MFGOracle, KerasAdam, pricing_loss, mfg_loss
        from oracle.pricing import sample_pricing_noise
        from oracle.mfg import sample_mfg_noise
        gen = torch.Generator().manual_seed(0)

        def cpu_ms(make_it):
            it = make_it()
            it()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); it(); ts.append(time.perf_counter() - t0)
            return 1e3 * float(np.median(ts))

        def pricing_it(kind, model, scheme, B, Mc):
            layout = H.pricing_layout(kind, scheme, 1)
            theta = torch.tensor(H.random_theta(layout, 0), requires_grad=True)
            opt = KerasAdam(layout.total, 3e-4)

            def it():
                noise = sample_pricing_noise(model, scheme, B, max(Mc, 1), gen)
                theta.grad = None
                loss = pricing_loss(model, scheme, layout, theta, noise, B)
                loss.backward()
                opt.step(theta.data, theta.grad)
            return it

        def mfg_it(model, scheme, B):
            layout = H.mfg_layout(scheme)
            theta = torch.tensor(H.random_theta(layout, 0), requires_grad=True)
            opt = KerasAdam(layout.total, 1e-3)

            def it():
                noise = sample_mfg_noise(model, B, gen)
                theta.grad = None
                lh, li = mfg_loss(model, scheme, layout, theta, noise, B)
                (lh + li).backward()
                opt.step(theta.data, theta.grad)
            return it

        scheme_of = {"SolverGlobalFBSDE": "Global", "SolverMultiStepFBSDE1": "MultiStep1", "SolverMultiStepFBSDE2": "MultiStep2",
                     "SolverSumLocalFBSDE1": "SumLocal1", "SolverSumLocalFBSDE2": "SumLocal2", "SolverGlobalSumLocalReg": "SumLocalReg",
                     "SolverGlobalMultiStepReg": "MultiStepReg", "SolverMultiStepFBSDE": "MultiStep", "SolverSumLocalFBSDE": "SumLocal"}
        om = MertonOracle(aLin=0.1, limit=30, d=1, **H.MERTON)
        ov = VGOracle(aLin=0.1, **H.VG)
        og = MFGOracle(**P)
        for o in out:
            sch = scheme_of[o["solver"]]
            if o["config"].startswith("1"):
                ms = cpu_ms(lambda: pricing_it("merton", om, sch, o["paths"], o["M"]))
            elif o["config"].startswith("2"):
                ms = cpu_ms(lambda: pricing_it("vg", ov, sch, o["paths"], o["M"]))
            else:
                ms = cpu_ms(lambda: mfg_it(og, sch, o["paths"]))
            o["cpu_ms_per_iter"], o["cpu_threads"], o["speedup_vs_cpu"] = ms, torch.get_num_threads(), ms / o["ms_per_iter"]
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()

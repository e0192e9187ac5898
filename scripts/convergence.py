#!/usr/bin/env python
"""Convergence to the reference's known answers (with Y == A the coupling vanishes, so the exact Y0 of the coupled
FBSDE is the closed-form price - mainMerton.py:68-73, mainVG.py:65-70 plot the learned Y0 against it).

Trains, through the drop-in classes on one B200,
  merton_global   SolverGlobalFBSDE at the mainMerton.py defaults (B = 10, N = 50, M = 5000, lr 4e-4, 120 x 100 steps)  -> 0.2714569
  merton_reg      SolverGlobalSumLocalReg at the mainMerton.py defaults (train batch 1000 x 10, lr 3e-4, 120 x 100)       -> 0.2714569
  basket_d10      SolverGlobalSumLocalReg, Merton d = 10 geometric basket, N = 100, 66 000 paths per step (~2^16)           -> 0.1109224
  vg_global       SolverGlobalFBSDE (pure jump) at the mainVG.py defaults (B = 10, N = 30, M = 5000, lr 5e-4, 120 x 100)  -> 0.1331402
and compares the learned Y0 with the closed form in units of the standard error of a plain Monte-Carlo price estimate
(discounted payoff of the uncoupled model, simulated here with NumPy):
  se_batch   from as many paths as ONE training batch          se_epoch   from the paths of one outer epoch (num_epoch batches)
Writes profiles/r2_convergence_<case>.csv (epoch, Y0, validation loss, seconds) and prints one JSON line per case.
"""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def payoff_std_merton(p, d, n=400000, seed=1):
    """Standard deviation of e^{-rT} g(X_T) under the uncoupled Merton dynamics (exact in law: one step to T)."""
    rng = np.random.default_rng(seed)
    T, r, sig, lam, muJ, sJ, K, x0 = p["T"], p["r"], p["sigma"], p["lam"], p["muJ"], p["sigmaJ"], p["K"], p["x0"]
    drift = r - 0.5 * sig ** 2 - lam * (np.exp(muJ + 0.5 * sJ ** 2) - 1.0)
    dN = rng.poisson(lam * T, size=(n, d))
    logX = np.log(x0) + drift * T + sig * np.sqrt(T) * rng.standard_normal((n, d)) + dN * muJ + sJ * np.sqrt(dN) * rng.standard_normal((n, d))
    G = np.exp(logX.mean(axis=1))
    pay = np.exp(-r * T) * np.maximum(G - K, 0.0)
    return float(pay.std()), float(pay.mean()), float(pay.std() / np.sqrt(n))


def payoff_std_vg(p, n=400000, seed=2):
    rng = np.random.default_rng(seed)
    T, r, th, ka, sJ, K, x0 = p["T"], p["r"], p["theta"], p["kappa"], p["sigmaJ"], p["K"], p["x0"]
    corr = -np.log(1.0 - th * ka - 0.5 * ka * sJ ** 2) / ka
    g = rng.gamma(T / ka, ka, size=n)
    X = x0 * np.exp((r - corr) * T + th * g + sJ * np.sqrt(g) * rng.standard_normal(n))
    pay = np.exp(-r * T) * np.maximum(X - K, 0.0)
    return float(pay.std()), float(pay.mean()), float(pay.std() / np.sqrt(n))


def run_case(name, quick, ctx, write=True):
    import helpers as H
    from deepfbsdejsolvers_b200 import set_seed
    from deepfbsdejsolvers_b200 import coupledPricing as cp
    from deepfbsdejsolvers_b200.coupledPricing import SolversPureJump as pj
    set_seed(2026)
    M, V = H.MERTON, H.VG

    def net(bY0, nout):
        return cp.Net(bY0, nout, [21, 21], "tanh")
    if name in ("merton_global", "merton_reg"):
        mm = cp.MertonJumpModel(M["T"], M["N"], M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(0.1), 30)
        sd, mc, mc_se = payoff_std_merton(M, 1)
        if name == "merton_global":
            solver, batch = cp.SolverGlobalFBSDE(mm, net(1, 1), net(0, 1), 4e-4, ctx=ctx), 10
        else:
            solver, batch = cp.SolverGlobalSumLocalReg(mm, net(0, 1), net(0, 1), 3e-4, ctx=ctx), 10000
        args = (10, 100, 100, 30 if quick else 120)
    elif name == "basket_d10":
        P = dict(M, N=100)
        mm = cp.MertonJumpModel(P["T"], P["N"], P["r"], P["muJ"], P["sigmaJ"], P["sigma"], P["lam"], P["K"], P["x0"], cp.AbsCoupling(0.1), 100, d=10)
        sd, mc, mc_se = payoff_std_merton(P, 10)
        solver, batch = cp.SolverGlobalSumLocalReg(mm, net(0, 1), net(0, 1), 1e-3, ctx=ctx), 66000
        args = (66, 10, 100, 15 if quick else 60)
    else:
        mm = cp.VGmodel(V["T"], V["N"], V["r"], V["theta"], V["kappa"], V["sigmaJ"], V["K"], V["x0"], cp.AbsCoupling(0.1))
        sd, mc, mc_se = payoff_std_vg(V)
        solver, batch = pj.SolverGlobalFBSDE(mm, net(0, 1), net(1, 1), 5e-4, ctx=ctx), 10
        args = (10, 100, 100, 30 if quick else 120)
    closed = float(mm.A(0, mm.init(1)).numpy()[0])
    with contextlib.redirect_stdout(io.StringIO()):
        solver.train(*args)
    y0 = np.array(solver.listY0, dtype=np.float64)
    tail = y0[-max(3, len(y0) // 10):]                 # the last tenth of the outer epochs (the reference plots the whole curve)
    se_batch, se_epoch = sd / np.sqrt(batch), sd / np.sqrt(batch * args[2])
    out = {"case": name, "solver": type(solver).__name__, "train_batch": batch, "steps": args[2] * args[3], "closed_form": closed,
           "mc_price": mc, "mc_price_se": mc_se, "Y0_last": float(y0[-1]), "Y0_tail_mean": float(tail.mean()),
           "abs_err_last": abs(float(y0[-1]) - closed), "abs_err_tail_mean": abs(float(tail.mean()) - closed),
           "payoff_std": sd, "se_batch": se_batch, "se_epoch": se_epoch,
           "err_in_se_batch": abs(float(tail.mean()) - closed) / se_batch, "err_in_se_epoch": abs(float(tail.mean()) - closed) / se_epoch,
           "train_seconds": float(solver.duration), "final_validation_loss": float(solver.lossList[-1])}
    if not write:
        return out
    path = os.path.join(ROOT, "profiles", "r2_convergence_%s.csv" % name)
    with open(path, "w") as f:
        f.write("epoch,Y0,validation_loss,seconds,closed_form\n")
        for k, (y, l, t) in enumerate(zip(solver.listY0, solver.lossList, solver.durationList)):
            f.write("%d,%.8f,%.6e,%.3f,%.8f\n" % (k, float(y), float(l), float(t), closed))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="merton_global,merton_reg,basket_d10,vg_global")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    from deepfbsdejsolvers_b200 import Context
    ctx = Context.default(0)
    for name in a.cases.split(","):
        print(json.dumps(run_case(name, a.quick, ctx)), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES, unmodified, from /root/reference.

TensorFlow is not installable in this image, so the files are imported against tests/golden/tfshim (a torch-backed
stand-in for the TF API they use).  For every solver class one real `train(...)` call with one Adam step is run;
the fixture stores the initial parameters (library flat layout), the random increments the reference drew (recorded at
its `tf.random.*` / `mathModel.jumps` / `mathModel.dN` call sites), and what the reference computed: the loss, the
tape gradients, the parameters after `optimizer.apply_gradients`, and the Y0 it reports.

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_golden_gpu.py) are both checked against these.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tfshim"))
REF = os.environ.get("FBSDEJ_REFERENCE", "/root/reference")

import tensorflow as tf  # noqa: E402  (the shim)


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


PM = load("ref_pricingModels", "coupledPricing/pricingModels.py")
NETP = load("ref_Networks_pricing", "coupledPricing/Networks.py")
SJD = load("ref_SolversJumpDiff", "coupledPricing/SolversJumpDiff.py")
SPJ = load("ref_SolversPureJump", "coupledPricing/SolversPureJump.py")
MFGM = load("ref_MFGModel", "coupledMFG/MFGModel.py")
NETM = load("ref_Networks_mfg", "coupledMFG/Networks.py")
MFGS = load("ref_MFGSolvers", "coupledMFG/MFGSolvers.py")

ALIN = 0.1
MCOMP = 5000   # hard-coded in the reference (SolversJumpDiff.py:34)


def func(x):   # mainMerton.py:60-61
    return ALIN * tf.math.abs(x)


def build_net(net, example):
    """Run the Keras-style lazy build, then give the biases non-zero values (a choice of parameter values)."""
    net(example)
    g = torch.Generator().manual_seed(99)
    for lyr in net.ListOfDense:
        lyr.bias.data = 0.1 * torch.randn(lyr.bias.shape, generator=g)


def flat_net(net):
    return np.concatenate([np.concatenate([l.kernel.detach().numpy().reshape(-1), l.bias.detach().numpy()]) for l in net.ListOfDense])


def flat_grads(net, var_list, grads):
    """Gradients of `net`'s Dense variables in library order (zeros where the tape returned None / net not trained)."""
    lookup = {id(v): g for v, g in zip(var_list, grads)}
    parts = []
    for l in net.ListOfDense:
        for v in (l.kernel, l.bias):
            g = lookup.get(id(v))
            parts.append(np.zeros(v.numel(), dtype=np.float32) if g is None else g.numpy().reshape(-1))
    return np.concatenate(parts)


class Recorder:
    """Records the increments at the reference's own call sites."""

    def __init__(self, model, B):
        self.B, self.inside, self.J, self.JMC, self.gauss = B, False, [], [], []
        orig_jumps, orig_normal, rec = model.jumps, tf.random.normal, self

        def jumps(n):
            rec.inside = True
            out = orig_jumps(n)
            rec.inside = False
            (rec.J if n == rec.B else rec.JMC).append(out.detach().clone())
            return out

        def normal(shape, *a, **k):
            x = orig_normal(shape, *a, **k)
            if not rec.inside:
                rec.gauss.append(x.detach().clone())
            return x
        model.jumps = jumps
        tf.random.normal = normal
        self._restore = lambda: setattr(tf.random, "normal", orig_normal)

    def close(self):
        self._restore()


def pricing_case(kind, scheme):
    torch.manual_seed(0)
    tf.random.seed(1000 + zlib.crc32(f"{kind}/{scheme}".encode()) % 1000)   # (hash() of a str is salted per process)
    tf.keras.initializers.GEN.manual_seed(7 + len(scheme))
    tf.GradientTape.LOG.clear()
    merton = kind == "merton"
    B = 6
    if merton:
        N = 6
        par = dict(T=1.0, N=N, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)
        model = PM.MertonJumpModel(par["T"], N, par["r"], par["muJ"], par["sigmaJ"], par["sigma"], par["lam"], par["K"], par["x0"], func, 30)
        S = SJD
    else:
        N = 4
        par = dict(T=1.0, N=N, r=0.1, theta=-0.1, kappa=0.1, sigmaJ=0.2, K=1.0, x0=1.0)
        model = PM.VGmodel(par["T"], N, par["r"], par["theta"], par["kappa"], par["sigmaJ"], par["K"], par["x0"], func)
        S = SPJ
    reg = scheme.endswith("Reg")
    one = scheme.endswith("1")
    y0_on = None
    if scheme == "Global":
        y0_on = "UZ" if merton else "Gam"
    ndim = 1 if (scheme == "Global" or reg or not merton) else 2
    layer = 21 * np.ones((2,), dtype=np.int32)
    netA = NETP.Net(1 if y0_on == "UZ" else 0, ndim, layer, "tanh")
    netB = None if one else NETP.Net(1 if y0_on == "Gam" else 0, 1, layer, "tanh")
    build_net(netA, tf.zeros([1, 2]))
    if netB is not None:
        build_net(netB, tf.zeros([1, 3]))
    lr = 3e-4
    cls = {"Global": "SolverGlobalFBSDE", "MultiStep1": "SolverMultiStepFBSDE1", "MultiStep2": "SolverMultiStepFBSDE2",
           "SumLocal1": "SolverSumLocalFBSDE1", "SumLocal2": "SolverSumLocalFBSDE2", "SumLocalReg": "SolverGlobalSumLocalReg",
           "MultiStepReg": "SolverGlobalMultiStepReg"}[scheme]
    solver = getattr(S, cls)(model, netA, lr) if one else getattr(S, cls)(model, netA, netB, lr)
    theta0 = [flat_net(netA)] + ([flat_net(netB)] if netB is not None else [])
    if y0_on:
        theta0.append(np.array([(netA if y0_on == "UZ" else netB).Y0.detach().numpy()], dtype=np.float32))
    theta0 = np.concatenate(theta0).astype(np.float32)
    # the Reg solvers train on 1000*batchSize paths (SolversJumpDiff.py:435): batchSize=1 keeps the fixture small
    bs = 1 if reg else B
    Btrain = 1000 * bs if reg else bs
    rec = Recorder(model, Btrain)
    solver.train(bs, 1, 1, 1)     # one Adam step, then a 1-path (x100 for Reg) validation pass we do not record
    rec.close()
    loss, grads, trained = tf.GradientTape.LOG[0]     # the variable list the reference handed to tape.gradient
    g = [flat_grads(netA, trained, grads)] + ([flat_grads(netB, trained, grads)] if netB is not None else [])
    if y0_on:
        holder = netA if y0_on == "UZ" else netB
        lookup = {id(v): gg for v, gg in zip(trained, grads)}
        g.append(np.array([lookup[id(holder.Y0)].numpy()], dtype=np.float32))
    theta1 = [flat_net(netA)] + ([flat_net(netB)] if netB is not None else [])
    if y0_on:
        theta1.append(np.array([(netA if y0_on == "UZ" else netB).Y0.detach().numpy()], dtype=np.float32))
    ncall = N + 1 if scheme.startswith("SumLocal") and not reg else N      # jumps() calls of the training pass
    out = dict(kind=kind, scheme=scheme, B=Btrain, N=N, lr=lr, theta0=theta0, loss=np.float64(loss),
               grad=np.concatenate(g).astype(np.float32), theta1=np.concatenate(theta1).astype(np.float32),
               Y0_report=np.float32(solver.listY0[0]), J=torch.stack(rec.J[:ncall][:N], 0).numpy(),
               **{k: np.float64(v) for k, v in par.items() if k != "N"})
    if merton:
        out["dW"] = (np.float32(np.sqrt(model.dt)) * torch.stack(rec.gauss[:N], 0)).numpy()
    if not reg:
        out["JMC"] = torch.stack(rec.JMC[:ncall][:N], 0).numpy()
    return out


def trajectory_case(nsteps=25):
    """A multi-step training trajectory of the reference's own code: SolverGlobalFBSDE (Merton) for `nsteps` Adam steps.  The fixture
    holds the increments of every step and, per step, the reference's loss and its trainable Y0 after the update - what the
    north star calls "the reference's own price" on the way."""
    torch.manual_seed(0)
    tf.random.seed(4242)
    tf.keras.initializers.GEN.manual_seed(31)
    tf.GradientTape.LOG.clear()
    B, N = 8, 4
    par = dict(T=1.0, N=N, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)
    model = PM.MertonJumpModel(par["T"], N, par["r"], par["muJ"], par["sigmaJ"], par["sigma"], par["lam"], par["K"], par["x0"], func, 30)
    layer = 21 * np.ones((2,), dtype=np.int32)
    netA, netB = NETP.Net(1, 1, layer, "tanh"), NETP.Net(0, 1, layer, "tanh")
    build_net(netA, tf.zeros([1, 2]))
    build_net(netB, tf.zeros([1, 3]))
    lr = 4e-4                                           # mainMerton.py:18
    solver = SJD.SolverGlobalFBSDE(model, netA, netB, lr)
    theta0 = np.concatenate([flat_net(netA), flat_net(netB), np.array([netA.Y0.detach().numpy()], dtype=np.float32)]).astype(np.float32)
    y0_after = []
    orig_apply = tf.keras.optimizers.Adam.apply_gradients

    def apply(self, gv):
        orig_apply(self, gv)
        y0_after.append(float(netA.Y0.detach()))
    tf.keras.optimizers.Adam.apply_gradients = apply
    rec = Recorder(model, B)
    solver.train(B, 1, nsteps, 1)
    rec.close()
    tf.keras.optimizers.Adam.apply_gradients = orig_apply
    losses = np.array([tf.GradientTape.LOG[k][0] for k in range(nsteps)], dtype=np.float64)
    theta1 = np.concatenate([flat_net(netA), flat_net(netB), np.array([netA.Y0.detach().numpy()], dtype=np.float32)]).astype(np.float32)
    sq = np.float32(np.sqrt(model.dt))
    k = nsteps * N
    return dict(kind="merton", scheme="Global", B=B, N=N, lr=lr, nsteps=nsteps, theta0=theta0, theta_final=theta1, losses=losses,
                Y0_after_step=np.array(y0_after[:nsteps], dtype=np.float32),
                dW=(sq * torch.stack(rec.gauss[:k], 0)).numpy().reshape(nsteps, N, B),
                J=torch.stack(rec.J[:k], 0).numpy().reshape(nsteps, N, B),
                JMC=torch.stack(rec.JMC[:k], 0).numpy().reshape(nsteps, N, MCOMP),
                **{kk: np.float64(v) for kk, v in par.items() if kk != "N"})


def reg_trajectory_case(kind="merton", nsteps=300, seed=20261018):
    """A LONG training trajectory of the reference's own SolverGlobalSumLocalReg (the headline scheme; d = 1, 1000 paths per step as
    its train() draws them - SolversJumpDiff.py:435, SolversPureJump.py:403): `nsteps` consecutive Adam steps.  The increments are
    INJECTED at the reference's draw sites (tf.random.normal([nbSimul]) and mathModel.jumps(nbSimul), SolversJumpDiff.py:402-405,
    SolversPureJump.py:372) from NumPy's frozen RandomState stream (tests/golden/noise_streams.py), so the fixture holds the seed,
    the reference's loss at every step and its reported Y0 = U(0, x0) after every update - not the arrays."""
    import noise_streams as TH
    merton = kind == "merton"
    torch.manual_seed(0)
    tf.random.seed(777)
    tf.keras.initializers.GEN.manual_seed(41 if merton else 43)
    tf.GradientTape.LOG.clear()
    B = 1000
    if merton:
        N = 6
        par = dict(T=1.0, N=N, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)
        model = PM.MertonJumpModel(par["T"], N, par["r"], par["muJ"], par["sigmaJ"], par["sigma"], par["lam"], par["K"], par["x0"], func, 30)
        S, lr = SJD, 3e-4                               # mainMerton.py:20
        dW, J = TH.reg_trajectory_noise(seed, nsteps, N, B, model.dt, par["lam"], par["muJ"], par["sigmaJ"])
    else:
        N = 5
        par = dict(T=1.0, N=N, r=0.1, theta=-0.1, kappa=0.1, sigmaJ=0.2, K=1.0, x0=1.0)
        model = PM.VGmodel(par["T"], N, par["r"], par["theta"], par["kappa"], par["sigmaJ"], par["K"], par["x0"], func)
        S, lr = SPJ, 1.5e-4                             # mainVG.py:20
        dW, J = None, TH.vg_trajectory_noise(seed + 1, nsteps, N, B, model.dt, par["theta"], par["kappa"], par["sigmaJ"])
    layer = 21 * np.ones((2,), dtype=np.int32)
    netA, netB = NETP.Net(0, 1, layer, "tanh"), NETP.Net(0, 1, layer, "tanh")
    build_net(netA, tf.zeros([1, 2]))
    build_net(netB, tf.zeros([1, 3]))
    solver = S.SolverGlobalSumLocalReg(model, netA, netB, lr)
    theta0 = np.concatenate([flat_net(netA), flat_net(netB)]).astype(np.float32)
    sq = np.float32(np.sqrt(model.dt))
    cnt = {"g": 0, "j": 0}
    orig_jumps, orig_normal = model.jumps, tf.random.normal

    def jumps(n):
        if n != B:
            return orig_jumps(n)                        # the validation pass (100 paths)
        k = cnt["j"]
        cnt["j"] += 1
        return torch.tensor(J[k // N, k % N])

    def normal(shape, *a, **kw):
        if dW is None or list(shape) != [B]:
            return orig_normal(shape, *a, **kw)
        k = cnt["g"]
        cnt["g"] += 1
        return torch.tensor(dW[k // N, k % N] / sq)     # the reference multiplies by sqrt(dt) itself
    model.jumps, tf.random.normal = jumps, normal
    y0_after = []
    orig_apply = tf.keras.optimizers.Adam.apply_gradients

    def apply(self, gv):
        orig_apply(self, gv)
        with torch.no_grad():
            y0_after.append(float(netA(torch.tensor([[0.0, par["x0"]]], dtype=torch.float32))[0].reshape(-1)[0]))
    tf.keras.optimizers.Adam.apply_gradients = apply
    try:
        solver.train(1, 1, nsteps, 1)
    finally:
        tf.keras.optimizers.Adam.apply_gradients = orig_apply
        tf.random.normal = orig_normal
    assert cnt["j"] == nsteps * N and cnt["g"] == (nsteps * N if merton else 0), cnt
    losses = np.array([tf.GradientTape.LOG[k][0] for k in range(nsteps)], dtype=np.float64)
    theta1 = np.concatenate([flat_net(netA), flat_net(netB)]).astype(np.float32)
    out = dict(kind=kind, scheme="SumLocalReg", B=B, N=N, lr=lr, nsteps=nsteps, seed=seed, theta0=theta0, theta_final=theta1,
               losses=losses, Y0_after_step=np.array(y0_after[:nsteps], dtype=np.float32), Y0_report=np.float32(solver.listY0[0]),
               J_checksum=np.float64(J.astype(np.float64).sum()), **{kk: np.float64(v) for kk, v in par.items() if kk != "N"})
    if merton:
        out["dW_checksum"] = np.float64(dW.astype(np.float64).sum())
    return out


def jump_trajectory_case(kind="merton", nsteps=100, seed=20261019):
    """A long training trajectory of the reference's own SolverGlobalFBSDE at the DEFAULT shapes of mainMerton.py / mainVG.py
    (10 paths, N = 50 / 30 time steps, 5000 compensator samples redrawn at every time step, lr 4e-4 / 5e-4): `nsteps` consecutive
    Adam steps on increments injected at the reference's draw sites (SolversJumpDiff.py:30-34, SolversPureJump.py:31-32) from the
    frozen RandomState stream; the fixture holds the seed, the reference's losses and its trainable Y0 after every update."""
    import noise_streams as TH
    merton = kind == "merton"
    torch.manual_seed(0)
    tf.random.seed(555)
    tf.keras.initializers.GEN.manual_seed(53 if merton else 59)
    tf.GradientTape.LOG.clear()
    B = 10
    if merton:
        N = 50
        par = dict(T=1.0, N=N, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)
        model = PM.MertonJumpModel(par["T"], N, par["r"], par["muJ"], par["sigmaJ"], par["sigma"], par["lam"], par["K"], par["x0"], func, 30)
        S, lr, y0_on = SJD, 4e-4, "UZ"                  # mainMerton.py:18
    else:
        N = 30
        par = dict(T=1.0, N=N, r=0.1, theta=-0.1, kappa=0.1, sigmaJ=0.2, K=1.0, x0=1.0)
        model = PM.VGmodel(par["T"], N, par["r"], par["theta"], par["kappa"], par["sigmaJ"], par["K"], par["x0"], func)
        S, lr, y0_on = SPJ, 5e-4, "Gam"                 # mainVG.py:18
    dW, J, JMC = TH.jump_trajectory_noise(kind, seed, nsteps, N, B, MCOMP, model.dt, par)
    layer = 21 * np.ones((2,), dtype=np.int32)
    netA = NETP.Net(1 if y0_on == "UZ" else 0, 1, layer, "tanh")
    netB = NETP.Net(1 if y0_on == "Gam" else 0, 1, layer, "tanh")
    build_net(netA, tf.zeros([1, 2]))
    build_net(netB, tf.zeros([1, 3]))
    holder = netA if y0_on == "UZ" else netB
    solver = S.SolverGlobalFBSDE(model, netA, netB, lr)

    def theta():
        return np.concatenate([flat_net(netA), flat_net(netB), np.array([holder.Y0.detach().numpy()], dtype=np.float32)]).astype(np.float32)
    theta0 = theta()
    sq = np.float32(np.sqrt(model.dt))
    cnt = {"g": 0, "j": 0, "m": 0}
    tot = nsteps * N
    orig_jumps, orig_normal, orig_apply = model.jumps, tf.random.normal, tf.keras.optimizers.Adam.apply_gradients

    def jumps(n):
        key, arr = ("j", J) if n == B else ("m", JMC)
        k = cnt[key]
        if (n != B and n != MCOMP) or k >= tot:
            return orig_jumps(n)                        # the validation pass
        cnt[key] += 1
        return torch.tensor(arr[k // N, k % N])

    def normal(shape, *a, **kw):
        if dW is None or list(shape) != [B] or cnt["g"] >= tot:
            return orig_normal(shape, *a, **kw)
        k = cnt["g"]
        cnt["g"] += 1
        return torch.tensor(dW[k // N, k % N] / sq)
    y0_after = []

    def apply(self, gv):
        orig_apply(self, gv)
        y0_after.append(float(holder.Y0.detach()))
    model.jumps, tf.random.normal, tf.keras.optimizers.Adam.apply_gradients = jumps, normal, apply
    try:
        solver.train(B, 1, nsteps, 1)
    finally:
        tf.random.normal, tf.keras.optimizers.Adam.apply_gradients = orig_normal, orig_apply
    assert cnt["j"] == tot and cnt["m"] == tot and cnt["g"] == (tot if merton else 0), cnt
    losses = np.array([tf.GradientTape.LOG[k][0] for k in range(nsteps)], dtype=np.float64)
    out = dict(kind=kind, scheme="Global", B=B, N=N, M=MCOMP, lr=lr, nsteps=nsteps, seed=seed, theta0=theta0, theta_final=theta(),
               losses=losses, Y0_after_step=np.array(y0_after[:nsteps], dtype=np.float32),
               J_checksum=np.float64(J.astype(np.float64).sum()), JMC_checksum=np.float64(JMC.astype(np.float64).sum()),
               **{kk: np.float64(v) for kk, v in par.items() if kk != "N"})
    if merton:
        out["dW_checksum"] = np.float64(dW.astype(np.float64).sum())
    return out


def qaver_curve():
    t = np.arange(48) / 48.0
    return (0.35 + 0.2 * np.sin(2 * np.pi * (t - 0.3)) + 0.05 * np.sin(4 * np.pi * t))[:13]    # N = 12 steps


def mfg_case(scheme, couplage="ON"):
    tf.random.seed(2000 + len(scheme))
    tf.keras.initializers.GEN.manual_seed(17 + len(scheme))
    tf.GradientTape.LOG.clear()
    Q = qaver_curve()
    par = dict(T=0.25, R0=0.24, jumpFactor=8.0, alpha=30.0, beta=float(np.exp(-15.0)), coeffOU=5.0, A=150.0, K=50.0, pi=0.1,
               p0=6.159423723, p1=87.4286117, f0=0.0, f1=1e4, theta=0.12, C=80.0, S0=0.0, h1=0.0, h2=600.0, sig0=0.1, sig=0.3,
               alphaTarget=-0.2, coeffEqui=1.0)
    Qt = torch.tensor(Q, dtype=torch.float32)
    MFGM.QAver = Qt     # the reference reads a bare global `QAver` at MFGModel.py:67-68 (notebook namespace)
    model = MFGM.ModelCoupledFBSDE(par["T"], Qt, par["R0"], par["jumpFactor"], par["alpha"], par["beta"], par["coeffOU"], par["A"],
                                   par["K"], par["pi"], par["p0"], par["p1"], par["f0"], par["f1"], par["theta"], par["C"], par["S0"],
                                   par["h1"], par["h2"], par["sig0"], par["sig"], par["alphaTarget"], "stochastic", par["coeffEqui"])
    method = {"Global": "Global", "MultiStep": "SumMultiStep", "SumLocal": "SumLocal", "SumLocalReg": "SumLocalReg",
              "MultiStepReg": "SumMultiStepReg"}[scheme]
    reg = scheme.endswith("Reg")
    wh, wi = (2, 3) if scheme == "Global" else ((1, 1) if reg else (3, 4))
    km = NETM.kerasModels(NETM.Net_hat, NETM.Net, method, wh, wi, 20 * np.ones((2,), dtype=np.int32), 22 * np.ones((2,), dtype=np.int32),
                          "tanh", "tanh")
    model.init(1)
    build_net(km.model_hat, model.getProjectedStates())
    build_net(km.model, model.getAllStates())
    B, lr = 16, 1e-3
    cls = {"Global": "SolverGlobalFBSDE", "MultiStep": "SolverMultiStepFBSDE", "SumLocal": "SolverSumLocalFBSDE",
           "SumLocalReg": "SolverGlobalSumLocalReg", "MultiStepReg": "SolverGlobalMultiStepReg"}[scheme]
    solver = getattr(MFGS, cls)(model, km, lr, couplage)

    def theta():
        parts = [flat_net(km.model_hat), flat_net(km.model)]
        if scheme == "Global":
            parts.append(np.array([km.model_hat.Y0_hat.detach().numpy(), km.model.Y0.detach().numpy()], dtype=np.float32))
        return np.concatenate(parts).astype(np.float32)
    theta0 = theta()
    gauss, dNs = [], []
    orig_normal, orig_dN = tf.random.normal, model.dN

    def normal(shape, *a, **k):
        x = orig_normal(shape, *a, **k)
        gauss.append(x.detach().clone())
        return x

    def dN():
        n, c = orig_dN()
        dNs.append(n.detach().clone())
        return n, c
    tf.random.normal, model.dN = normal, dN
    solver.train(B, 1, 1, 1)
    tf.random.normal = orig_normal
    if couplage == "OFF":
        # MFGSolvers.py:92-115: one Adam step on the projected player's loss (variables of model_hat), a validation pass,
        # then one step on the individual player's loss (variables of model) with the SAME optimizer object (t = 2)
        N, sq = model.N, np.float32(np.sqrt(model.dt))
        out = dict(kind="mfg", scheme=scheme, couplage="OFF", B=B, N=N, lr=lr, QAver=Q, theta0=theta0, theta2=theta(),
                   Y0_hat_report=np.float32(solver.listY0_hat[0]), Y0_report=np.float32(solver.listY0[0]),
                   **{k: np.float64(v) for k, v in par.items()})
        for ph, call in ((1, 0), (2, 2)):            # optimizeBSDE calls: train-hat, validation, train-ind, validation
            loss, grads, trained = tf.GradientTape.LOG[ph - 1]
            lookup = {id(v): g for v, g in zip(trained, grads)}
            g = [flat_grads(km.model_hat, trained, grads), flat_grads(km.model, trained, grads)]
            if scheme == "Global":
                gy = [lookup.get(id(km.model_hat.Y0_hat)), lookup.get(id(km.model.Y0))]
                g.append(np.array([0.0 if x is None else float(x.numpy()) for x in gy], dtype=np.float32))
            gs, ds = gauss[2 * N * call:2 * N * (call + 1)], dNs[N * call:N * (call + 1)]
            out.update({f"p{ph}_loss": np.float64(loss), f"p{ph}_grad": np.concatenate(g).astype(np.float32),
                        f"p{ph}_dW0": (sq * torch.stack(gs[0::2], 0)).numpy(), f"p{ph}_dW": (sq * torch.stack(gs[1::2], 0)).numpy(),
                        f"p{ph}_dN": torch.stack(ds, 0).numpy()})
        return out
    loss, grads, trained = tf.GradientTape.LOG[0]
    lookup = {id(v): g for v, g in zip(trained, grads)}
    g = [flat_grads(km.model_hat, trained, grads), flat_grads(km.model, trained, grads)]
    if scheme == "Global":
        g.append(np.array([lookup[id(km.model_hat.Y0_hat)].numpy(), lookup[id(km.model.Y0)].numpy()], dtype=np.float32))
    N = model.N
    sq = np.float32(np.sqrt(model.dt))
    return dict(kind="mfg", scheme=scheme, B=B, N=N, lr=lr, QAver=Q, theta0=theta0, loss=np.float64(loss),
                grad=np.concatenate(g).astype(np.float32), theta1=theta(),
                Y0_hat_report=np.float32(solver.listY0_hat[0]), Y0_report=np.float32(solver.listY0[0]),
                dW0=(sq * torch.stack(gauss[0:2 * N:2], 0)).numpy(), dW=(sq * torch.stack(gauss[1:2 * N:2], 0)).numpy(),
                dN=torch.stack(dNs[:N], 0).numpy(), **{k: np.float64(v) for k, v in par.items()})


def mfg_trajectory_case(scheme="Global", nsteps=200):
    """A LONG training trajectory of the reference's own MFG SolverGlobalFBSDE (couplage ON, stochastic jumps): `nsteps` consecutive
    Adam steps over both networks and the two trainable initial values.  The increments the reference drew are recorded at its
    call sites (the Cox jump counts depend on the state, so they cannot be pre-drawn); per step the fixture holds the reference's
    loss and (Y0_hat, Y0) after the update."""
    tf.random.seed(3100)
    tf.keras.initializers.GEN.manual_seed(29)
    tf.GradientTape.LOG.clear()
    Q = qaver_curve()
    par = dict(T=0.25, R0=0.24, jumpFactor=8.0, alpha=30.0, beta=float(np.exp(-15.0)), coeffOU=5.0, A=150.0, K=50.0, pi=0.1,
               p0=6.159423723, p1=87.4286117, f0=0.0, f1=1e4, theta=0.12, C=80.0, S0=0.0, h1=0.0, h2=600.0, sig0=0.1, sig=0.3,
               alphaTarget=-0.2, coeffEqui=1.0)
    Qt = torch.tensor(Q, dtype=torch.float32)
    MFGM.QAver = Qt
    model = MFGM.ModelCoupledFBSDE(par["T"], Qt, par["R0"], par["jumpFactor"], par["alpha"], par["beta"], par["coeffOU"], par["A"],
                                   par["K"], par["pi"], par["p0"], par["p1"], par["f0"], par["f1"], par["theta"], par["C"], par["S0"],
                                   par["h1"], par["h2"], par["sig0"], par["sig"], par["alphaTarget"], "stochastic", par["coeffEqui"])
    km = NETM.kerasModels(NETM.Net_hat, NETM.Net, "Global", 2, 3, 20 * np.ones((2,), dtype=np.int32), 22 * np.ones((2,), dtype=np.int32),
                          "tanh", "tanh")
    model.init(1)
    build_net(km.model_hat, model.getProjectedStates())
    build_net(km.model, model.getAllStates())
    B, lr = 16, 1e-3                                    # mainMFGComparison.py:24
    solver = MFGS.SolverGlobalFBSDE(model, km, lr, "ON")

    def theta():
        return np.concatenate([flat_net(km.model_hat), flat_net(km.model),
                               np.array([km.model_hat.Y0_hat.detach().numpy(), km.model.Y0.detach().numpy()], dtype=np.float32)]).astype(np.float32)
    theta0 = theta()
    gauss, dNs, y0s = [], [], []
    orig_normal, orig_dN, orig_apply = tf.random.normal, model.dN, tf.keras.optimizers.Adam.apply_gradients

    def normal(shape, *a, **k):
        x = orig_normal(shape, *a, **k)
        gauss.append(x.detach().clone())
        return x

    def dN():
        n, c = orig_dN()
        dNs.append(n.detach().clone())
        return n, c

    def apply(self, gv):
        orig_apply(self, gv)
        y0s.append([float(km.model_hat.Y0_hat.detach()), float(km.model.Y0.detach())])
    tf.random.normal, model.dN, tf.keras.optimizers.Adam.apply_gradients = normal, dN, apply
    try:
        solver.train(B, 1, nsteps, 1)
    finally:
        tf.random.normal, tf.keras.optimizers.Adam.apply_gradients = orig_normal, orig_apply
    N, sq = model.N, np.float32(np.sqrt(model.dt))
    k = nsteps * N
    losses = np.array([tf.GradientTape.LOG[i][0] for i in range(nsteps)], dtype=np.float64)
    return dict(kind="mfg", scheme=scheme, B=B, N=N, lr=lr, nsteps=nsteps, QAver=Q, theta0=theta0, theta_final=theta(), losses=losses,
                Y0_after_step=np.array(y0s[:nsteps], dtype=np.float32),
                dW0=(sq * torch.stack(gauss[0:2 * k:2], 0)).numpy().reshape(nsteps, N, B),
                dW=(sq * torch.stack(gauss[1:2 * k:2], 0)).numpy().reshape(nsteps, N, B),
                dN=torch.stack(dNs[:k], 0).numpy().reshape(nsteps, N, B), **{kk: np.float64(v) for kk, v in par.items()})


def diag_case(scheme, nb=48):
    """The reference's own MFG diagnostics (MFGSolvers.py:118-178, 436-459) on fresh networks: simulateGlobalErr and followS, with
    the increments they drew."""
    tf.random.seed(3000 + len(scheme))
    tf.keras.initializers.GEN.manual_seed(23 + len(scheme))
    Q = qaver_curve()
    par = dict(T=0.25, R0=0.24, jumpFactor=8.0, alpha=30.0, beta=float(np.exp(-15.0)), coeffOU=5.0, A=150.0, K=50.0, pi=0.1,
               p0=6.159423723, p1=87.4286117, f0=0.0, f1=1e4, theta=0.12, C=80.0, S0=0.0, h1=0.0, h2=600.0, sig0=0.1, sig=0.3,
               alphaTarget=-0.2, coeffEqui=1.0)
    Qt = torch.tensor(Q, dtype=torch.float32)
    MFGM.QAver = Qt
    model = MFGM.ModelCoupledFBSDE(par["T"], Qt, par["R0"], par["jumpFactor"], par["alpha"], par["beta"], par["coeffOU"], par["A"],
                                   par["K"], par["pi"], par["p0"], par["p1"], par["f0"], par["f1"], par["theta"], par["C"], par["S0"],
                                   par["h1"], par["h2"], par["sig0"], par["sig"], par["alphaTarget"], "stochastic", par["coeffEqui"])
    method = {"Global": "Global", "SumLocal": "SumLocal"}[scheme]
    wh, wi = (2, 3) if scheme == "Global" else (3, 4)
    km = NETM.kerasModels(NETM.Net_hat, NETM.Net, method, wh, wi, 20 * np.ones((2,), dtype=np.int32), 22 * np.ones((2,), dtype=np.int32),
                          "tanh", "tanh")
    model.init(1)
    build_net(km.model_hat, model.getProjectedStates())
    build_net(km.model, model.getAllStates())
    cls = {"Global": "SolverGlobalFBSDE", "SumLocal": "SolverSumLocalFBSDE"}[scheme]
    solver = getattr(MFGS, cls)(model, km, 1e-3, "ON")
    parts = [flat_net(km.model_hat), flat_net(km.model)]
    if scheme == "Global":
        parts.append(np.array([km.model_hat.Y0_hat.detach().numpy(), km.model.Y0.detach().numpy()], dtype=np.float32))
    theta0 = np.concatenate(parts).astype(np.float32)
    out = dict(kind="mfg", scheme=scheme, nb=nb, N=model.N, QAver=Q, theta0=theta0, **{k: np.float64(v) for k, v in par.items()})
    N, sq = model.N, np.float32(np.sqrt(model.dt))
    for which in (("err", "follow") if hasattr(solver, "followS") else ("err",)):     # followS exists on SolverGlobalFBSDE only
        gauss, dNs = [], []
        orig_normal, orig_dN = tf.random.normal, model.dN

        def normal(shape, *a, **k):
            x = orig_normal(shape, *a, **k)
            gauss.append(x.detach().clone())
            return x

        def dN():
            n, c = orig_dN()
            dNs.append(n.detach().clone())
            return n, c
        tf.random.normal, model.dN = normal, dN
        res = solver.simulateGlobalErr(nb) if which == "err" else solver.followS(nb)
        tf.random.normal, model.dN = orig_normal, orig_dN
        out[which + "_dW0"] = (sq * torch.stack(gauss[0:2 * N:2], 0)).numpy()
        out[which + "_dW"] = (sq * torch.stack(gauss[1:2 * N:2], 0)).numpy()
        out[which + "_dN"] = torch.stack(dNs[:N], 0).numpy()
        if which == "err":
            out["err_result"] = np.array([float(x) for x in res], dtype=np.float64)
        else:
            out["follow_result"] = np.array([[float(v) for v in col] for col in res], dtype=np.float64)     # [4][N+1]
    return out


def main():
    import contextlib
    import io
    for kind in ("merton", "vg"):
        for scheme in ("Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2", "SumLocalReg", "MultiStepReg"):
            with contextlib.redirect_stdout(io.StringIO()):
                d = pricing_case(kind, scheme)
            np.savez_compressed(os.path.join(HERE, f"{kind}_{scheme}.npz"), **d)
            print(kind, scheme, "loss", float(d["loss"]), "Y0", float(d["Y0_report"]), "|grad|max", float(np.abs(d["grad"]).max()))
    for scheme in ("Global", "MultiStep", "SumLocal", "SumLocalReg", "MultiStepReg"):
        with contextlib.redirect_stdout(io.StringIO()):
            d = mfg_case(scheme)
        np.savez_compressed(os.path.join(HERE, f"mfg_{scheme}.npz"), **d)
        print("mfg", scheme, "loss", float(d["loss"]), "Y0_hat", float(d["Y0_hat_report"]), "Y0", float(d["Y0_report"]))
    with contextlib.redirect_stdout(io.StringIO()):
        d = trajectory_case()
    np.savez_compressed(os.path.join(HERE, "traj", "merton_Global_25steps.npz"), **d)
    print("trajectory: merton Global", int(d["nsteps"]), "steps, loss", float(d["losses"][0]), "->", float(d["losses"][-1]),
          "Y0", float(d["Y0_after_step"][0]), "->", float(d["Y0_after_step"][-1]))
    for kind in ("merton", "vg"):
        with contextlib.redirect_stdout(io.StringIO()):
            d = reg_trajectory_case(kind)
        np.savez_compressed(os.path.join(HERE, "traj", f"{kind}_SumLocalReg_300steps.npz"), **d)
        print("trajectory:", kind, "SumLocalReg", int(d["nsteps"]), "steps, loss", float(d["losses"][0]), "->", float(d["losses"][-1]),
              "Y0", float(d["Y0_after_step"][0]), "->", float(d["Y0_after_step"][-1]))
    for kind in ("merton", "vg"):
        with contextlib.redirect_stdout(io.StringIO()):
            d = jump_trajectory_case(kind)
        np.savez_compressed(os.path.join(HERE, "traj", f"{kind}_Global_defaults_100steps.npz"), **d)
        print("trajectory:", kind, "Global at the reference's default shapes,", int(d["nsteps"]), "steps, loss", float(d["losses"][0]), "->",
              float(d["losses"][-1]), "Y0", float(d["Y0_after_step"][0]), "->", float(d["Y0_after_step"][-1]))
    with contextlib.redirect_stdout(io.StringIO()):
        d = mfg_trajectory_case()
    np.savez_compressed(os.path.join(HERE, "traj", "mfg_Global_200steps.npz"), **d)
    print("trajectory: mfg Global", int(d["nsteps"]), "steps, loss", float(d["losses"][0]), "->", float(d["losses"][-1]),
          "Y0_hat, Y0", d["Y0_after_step"][0], "->", d["Y0_after_step"][-1])
    for scheme in ("Global", "SumLocal"):
        with contextlib.redirect_stdout(io.StringIO()):
            d = diag_case(scheme)
        np.savez_compressed(os.path.join(HERE, "diag", f"mfg_{scheme}_diagnostics.npz"), **d)
        print("mfg diagnostics", scheme, "simulateGlobalErr", d["err_result"])
    for scheme in ("Global", "SumLocalReg"):      # couplage OFF (the other OFF variants crash in the reference: MFGSolvers.py:291,431)
        with contextlib.redirect_stdout(io.StringIO()):
            d = mfg_case(scheme, couplage="OFF")
        np.savez_compressed(os.path.join(HERE, "off", f"mfg_{scheme}_OFF.npz"), **d)
        print("mfg OFF", scheme, "losses", float(d["p1_loss"]), float(d["p2_loss"]), "Y0_hat", float(d["Y0_hat_report"]), "Y0", float(d["Y0_report"]))


if __name__ == "__main__":
    main()

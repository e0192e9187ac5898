"""Builders shared by the parity tests: the same (model, scheme, parameters, noise) on the oracle and on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from oracle import MLPSpec, ParamLayout, MertonOracle, VGOracle, MFGOracle, pricing_loss, mfg_loss
from oracle.nets import init_params

MERTON = dict(T=1.0, N=50, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)      # mainMerton.py:57
VG = dict(T=1.0, N=30, r=0.1, theta=-0.1, kappa=0.1, sigmaJ=0.2, K=1.0, x0=1.0)                # mainVG.py:54
ALIN = 0.1

PRICING_SCHEME_ID = {"Global": 0, "MultiStep1": 1, "MultiStep2": 2, "SumLocal1": 3, "SumLocal2": 4, "SumLocalReg": 5,
                     "MultiStepReg": 6}
MFG_SCHEME_ID = {"Global": 0, "MultiStep": 2, "SumLocal": 4, "SumLocalReg": 5, "MultiStepReg": 6}


def qaver_curve(nbDays=2):
    """A smooth synthetic 48-points-per-day load curve in the range of mainMFGComparison.py:83-90 (values ~0.1-0.6)."""
    t = np.arange(48 * nbDays) / 48.0
    return 0.35 + 0.2 * np.sin(2 * np.pi * (t - 0.3)) + 0.05 * np.sin(4 * np.pi * t)


def mfg_params(nbDays=2, jumpModel="stochastic"):
    Q = qaver_curve(nbDays)
    return dict(T=float(nbDays), QAver=Q, R0=0.24, jumpFactor=8.0, alpha=30.0, beta=float(np.exp(-15.0)), coeffOU=5.0, A=150.0,
                K=50.0, pi=0.1, p0=6.159423723, p1=87.4286117, f0=0.0, f1=1e4, theta=0.12, C=80.0, S0=0.0, h1=0.0, h2=600.0,
                sig0=0.1, sig=0.3, alphaTarget=-0.2, jumpModel=jumpModel, coeffEqui=1.0)


def pricing_layout(kind, scheme, d, H=21, act="tanh", L=2):
    brown = kind == "merton"
    one = scheme.endswith("1")
    reg = scheme.endswith("Reg")
    if scheme == "Global":
        noutA = d if brown else 1
    elif reg:
        noutA = 1
    else:
        noutA = 1 + d if brown else 1
    nets = [MLPSpec(1 + d, [H] * L, noutA, act)]
    if not one:
        nets.append(MLPSpec(1 + 2 * d, [H] * L, 1, act))
    return ParamLayout(nets, n_y0=1 if scheme == "Global" else 0)


def mfg_layout(scheme, Hh=20, H=22, act="tanh", L=2):
    reg = scheme.endswith("Reg")
    if scheme == "Global":
        a, b = 2, 3
    elif reg:
        a, b = 1, 1
    else:
        a, b = 3, 4
    return ParamLayout([MLPSpec(4, [Hh] * L, a, act), MLPSpec(6, [H] * L, b, act)], n_y0=2 if scheme == "Global" else 0)


def random_theta(layout, seed, scale_bias=0.1):
    rng = np.random.default_rng(seed)
    th = init_params(layout, rng)
    # Keras starts biases at 0; perturb them so that bias gradients / paths are exercised away from the symmetric point
    for k in range(len(layout.nets)):
        for (w0, w1, fi, fo, b0, b1) in layout.net_slices(k):
            th[b0:b1] = scale_bias * rng.standard_normal(b1 - b0).astype(np.float32)
    for j in range(layout.n_y0):
        th[layout.y0_offset + j] = np.float32(0.2 + 0.1 * j)
    return th


def merton_noise(model, B, M, seed, with_jmc=True):
    """dW [N,B,d], J [N,B,d], JMC [N,M,d] drawn the way pricingModels.py:57-61 / SolversJumpDiff.py:30-34 do."""
    g = torch.Generator().manual_seed(seed)
    N, d = model.N, model.d
    noise = {"dW": (np.float32(np.sqrt(model.dt)) * torch.randn(N, B, d, generator=g)).float(),
             "J": torch.stack([model.jumps(B, g) for _ in range(N)], 0)}
    if with_jmc:
        noise["JMC"] = torch.stack([model.jumps(M, g) for _ in range(N)], 0)
    return noise


def vg_noise(model, B, M, seed, with_jmc=True):
    g = torch.Generator().manual_seed(seed)
    N = model.N
    noise = {"J": torch.stack([model.jumps(B, g) for _ in range(N)], 0)}
    if with_jmc:
        noise["JMC"] = torch.stack([model.jumps(M, g) for _ in range(N)], 0)
    return noise


def to_planes(x):
    """[N, B, d] (oracle) -> [N, d, B] (library)."""
    return np.ascontiguousarray(x.numpy().transpose(0, 2, 1))


def oracle_pricing(model, scheme, layout, theta, noise, B, dtype=torch.float32, stale_time=True):
    th = torch.tensor(theta, dtype=dtype, requires_grad=True)
    nz = {k: v.to(dtype) for k, v in noise.items()}
    old = model.dtype
    model.dtype = dtype
    aux = {}
    loss = pricing_loss(model, scheme, layout, th, nz, B, stale_time=stale_time, aux=aux)
    loss.backward()
    model.dtype = old
    return float(loss.detach()), th.grad.double().numpy(), {k: v.double().numpy() for k, v in aux.items()}


def oracle_mfg(model, scheme, layout, theta, noise, B, dtype=torch.float32, w=(1.0, 1.0)):
    th = torch.tensor(theta, dtype=dtype, requires_grad=True)
    nz = {k: v.to(dtype) for k, v in noise.items()}
    old = model.dtype
    model.dtype = dtype
    aux = {}
    lh, li = mfg_loss(model, scheme, layout, th, nz, B, aux=aux)
    (w[0] * lh + w[1] * li).backward()
    model.dtype = old
    return (float(lh.detach()), float(li.detach())), th.grad.double().numpy(), {k: v.double().numpy() for k, v in aux.items()}


# ---- native side --------------------------------------------------------------------------------------------
def native_pricing(ctx, kind, params, scheme, layout, d=1, M=0, limit=30, stale_time=True, price_table=False, tensor_cores=False):
    from deepfbsdejsolvers_b200 import NetSpec
    from deepfbsdejsolvers_b200.coupledPricing import MertonJumpModel, VGmodel, AbsCoupling
    if kind == "merton":
        mm = MertonJumpModel(params["T"], params["N"], params["r"], params["muJ"], params["sigmaJ"], params["sigma"],
                             params["lam"], params["K"], params["x0"], AbsCoupling(ALIN), limit, d=d, price_table=price_table)
    else:
        mm = VGmodel(params["T"], params["N"], params["r"], params["theta"], params["kappa"], params["sigmaJ"], params["K"],
                     params["x0"], AbsCoupling(ALIN))
    nets = [NetSpec(n.nin, n.hidden[0], n.nout, n.activation, len(n.hidden)) for n in layout.nets]
    return mm.make_solver(PRICING_SCHEME_ID[scheme], nets, layout.n_y0, M, ctx=ctx, stale_time=stale_time, tensor_cores=tensor_cores)


def native_mfg(ctx, params, scheme, layout, tensor_cores=False):
    from deepfbsdejsolvers_b200 import NetSpec
    from deepfbsdejsolvers_b200.coupledMFG import ModelCoupledFBSDE
    mm = ModelCoupledFBSDE(**params)
    nets = [NetSpec(n.nin, n.hidden[0], n.nout, n.activation, len(n.hidden)) for n in layout.nets]
    return mm.make_solver(MFG_SCHEME_ID[scheme], nets, layout.n_y0, ctx=ctx, tensor_cores=tensor_cores)


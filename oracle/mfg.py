"""Oracle for the coupled-MFG half of the reference (smart-grid mean-field game, Cox jumps).

Restates, with injectable noise:
  * coupledMFG/MFGModel.py:4-106  (ModelCoupledFBSDE)  -> MFGOracle  (the bare global `QAver`
    at MFGModel.py:67-68 is read as `self.QAver`)
  * coupledMFG/MFGSolvers.py loss graphs: Global :24-47, MultiStep :187-224, SumLocal :328-364,
    SumLocalReg :469-505, MultiStepReg :615-651  -> mfg_loss

Noise inputs (`noise` dict), all [N,B]: dW0 (common), dW (idiosyncratic), both already
sqrt(dt)*N(0,1), and dN (Poisson counts of the Cox process; the reference samples them from
Poisson(lam(hQ_i)*dt) at MFGModel.py:47-54 - injected here so both sides see identical jumps,
the same hook the reference's own fixed-trajectory replay uses, MFGSolutions.py:23-31).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .nets import ParamLayout, mlp_forward

MFG_SCHEMES = ("Global", "MultiStep", "SumLocal", "SumLocalReg", "MultiStepReg")


class MFGOracle:
    def __init__(self, T, QAver, R0, jumpFactor, alpha, beta, coeffOU, A, K, pi, p0, p1, f0, f1, theta, C, S0,
                 h1, h2, sig0, sig, alphaTarget, jumpModel, coeffEqui, dtype=torch.float32):
        self.T, self.QAver, self.R0, self.jumpFactor = T, np.asarray(QAver, dtype=np.float64), R0, jumpFactor
        self.alpha, self.beta, self.coeffOU = alpha, beta, coeffOU
        self.A, self.K, self.pi, self.p0, self.p1, self.f0, self.f1, self.theta = A, K, pi, p0, p1, f0, f1, theta
        self.C, self.S0, self.h1, self.h2, self.sig0, self.sig = C, S0, h1, h2, sig0, sig
        self.alphaTarget, self.jumpModel, self.coeffEqui = alphaTarget, jumpModel, coeffEqui
        self.N = len(self.QAver) - 1
        self.dt = T / self.N
        self.dtype = dtype

    def mean_hq(self, i: int) -> float:
        """MFGModel.py:67-68: e^{-k i dt} Q0 + k sum_{j<i} Q_j e^{k (j-i) dt} dt ; i = 0 -> Q0."""
        if i == 0:
            return float(self.QAver[0])
        k, dt = self.coeffOU, self.dt
        j = np.arange(i)
        return float(math.exp(-k * i * dt) * self.QAver[0] + k * np.sum(self.QAver[:i] * np.exp(k * (j - i) * dt) * dt))

    def init(self, B):  # :35-43
        one = torch.ones(B, dtype=self.dtype)
        return dict(hQ=float(self.QAver[0]) * one, Q=float(self.QAver[0]) * one, R=self.R0 * one,
                    hS=self.S0 * one, S=self.S0 * one, i=0)

    def lam_dt(self, st):  # :47-54
        if self.jumpModel == "stochastic":
            lam = self.beta * (torch.exp(self.alpha * st["hQ"]) - 1)
        else:
            lam = self.jumpFactor * torch.ones_like(st["hQ"])
        return lam * self.dt

    def _ind(self, st):
        return torch.where(st["R"] <= self.theta, torch.ones_like(st["R"]), torch.zeros_like(st["R"]))

    def _alpha_target(self, st):  # :76-79
        if self.jumpModel == "stochastic":
            return self.alphaTarget * self.mean_hq(st["i"])
        return self.alphaTarget

    def calpha_hat(self, st, hY):  # :82-85
        c, ind = self.coeffEqui, self._ind(st)
        kTheta = self.A + (1 - self.pi) * c * self.p1 + self.K + c * self.f1 * ind
        return -(1 / kTheta) * (self.p0 + self.pi * self.p1 * st["hQ"] + ((1 - self.pi) * c * self.p1 + self.K) * st["hQ"] + hY
                                + (self.f0 + c * self.f1 * (st["hQ"] - self.mean_hq(st["i"]) - self._alpha_target(st))) * ind)

    def calpha(self, st, hY, Y):  # :87-89
        c, ind = self.coeffEqui, self._ind(st)
        ah = self.calpha_hat(st, hY)
        return -(1 / (self.A + self.K)) * (self.K * st["Q"] + self.p0 + self.pi * self.p1 * st["hQ"]
                                           + (1 - self.pi) * c * self.p1 * (st["hQ"] + ah) + Y
                                           + (self.f0 + c * self.f1 * (st["hQ"] - self.mean_hq(st["i"]) + ah - self._alpha_target(st))) * ind)

    def one_step(self, st, dW0, dW, dN, hY, Y):  # :58-71 (update order preserved)
        i1 = st["i"] + 1
        hS = st["hS"] + self.calpha_hat(st, hY) * self.dt
        S = st["S"] + self.calpha(st, hY, Y) * self.dt
        R = st["R"] + self.dt - torch.where(dN > 0, st["R"], torch.zeros_like(st["R"]))
        q = float(self.QAver[i1])
        hQ = st["hQ"] + self.coeffOU * (q - st["hQ"]) * self.dt + self.sig0 * dW0
        Q = st["Q"] + self.coeffOU * (q - st["Q"]) * self.dt + self.sig0 * dW0 + self.sig * dW
        return dict(hQ=hQ, Q=Q, R=R, hS=hS, S=S, i=i1)

    def f(self, U):  # :92-93
        return U * self.C

    def g(self, X):  # :97-98
        return self.h1 + self.h2 * X

    def proj_states(self, st):  # :102-103
        t = st["i"] * self.dt * torch.ones_like(st["hQ"])
        return torch.stack([t, st["hQ"], st["hS"], st["R"]], dim=-1)

    def all_states(self, st):  # :106-107
        t = st["i"] * self.dt * torch.ones_like(st["hQ"])
        return torch.stack([t, st["Q"], st["S"], st["hQ"], st["hS"], st["R"]], dim=-1)


def mfg_loss(model: MFGOracle, scheme: str, layout: ParamLayout, theta: torch.Tensor, noise: Dict[str, torch.Tensor],
             B: int, aux: Optional[dict] = None):
    """Returns (loss_hat, loss_ind); the coupled objective (couplage 'ON') is their sum.

    layout: net 0 = model_hat (4 inputs), net 1 = model (6 inputs); Global has two trainable
    scalars theta[y0_offset] = Y0_hat, theta[y0_offset+1] = Y0.
    Output columns (mainMFGComparison.py:119-124): Global hat->[hZ0,hGam], model->[Z0,Gam,Z];
    MultiStep/SumLocal hat->[hY,hZ0,hGam], model->[Y,Z0,Gam,Z]; Reg hat->[hY], model->[Y].
    """
    assert scheme in MFG_SCHEMES
    N, dt = model.N, model.dt
    dW0, dW, dN = noise["dW0"], noise["dW"], noise["dN"]
    st = model.init(B)
    reg = scheme in ("SumLocalReg", "MultiStepReg")
    tr = {"hS": [st["hS"]], "S": [st["S"]], "hY": [], "Y": []}

    def nets(st):
        return (mlp_forward(theta, layout, 0, model.proj_states(st)), mlp_forward(theta, layout, 1, model.all_states(st)))

    def increments(st, oh, o, i, c0):
        """(-dt f(hS) + hZ0 dW0 + hGam (dN - lam dt), same for the individual player + Z dW)."""
        comp = model.lam_dt(st)
        a_h = -dt * model.f(st["hS"]) + oh[:, c0] * dW0[i] + oh[:, c0 + 1] * (dN[i] - comp)
        a = -dt * model.f(st["S"]) + o[:, c0] * dW0[i] + o[:, c0 + 1] * (dN[i] - comp) + o[:, c0 + 2] * dW[i]
        return a_h, a

    if scheme == "Global":  # MFGSolvers.py:24-47
        hY = theta[layout.y0_offset] * torch.ones(B, dtype=theta.dtype)
        Y = theta[layout.y0_offset + 1] * torch.ones(B, dtype=theta.dtype)
        for i in range(N):
            oh, o = nets(st)
            a_h, a = increments(st, oh, o, i, 0)
            tr["hY"].append(hY); tr["Y"].append(Y)
            st = model.one_step(st, dW0[i], dW[i], dN[i], hY, Y)   # OLD hY, Y (note ii)
            hY, Y = hY + a_h, Y + a
            tr["hS"].append(st["hS"]); tr["S"].append(st["S"])
        tr["hY"].append(hY); tr["Y"].append(Y)
        lh = torch.mean((hY - model.g(st["hS"])) ** 2)
        li = torch.mean((Y - model.g(st["S"])) ** 2)
    elif scheme in ("MultiStep", "MultiStepReg"):  # :187-224, :615-651
        hys, ys, ahs, as_ = [], [], [], []
        for i in range(N):
            oh, o = nets(st)
            hY, Y = oh[:, 0], o[:, 0]
            if reg:
                a_h, a = -dt * model.f(st["hS"]), -dt * model.f(st["S"])
            else:
                a_h, a = increments(st, oh, o, i, 1)
            hys.append(hY); ys.append(Y); ahs.append(a_h); as_.append(a)
            tr["hY"].append(hY); tr["Y"].append(Y)
            st = model.one_step(st, dW0[i], dW[i], dN[i], hY, Y)
            tr["hS"].append(st["hS"]); tr["S"].append(st["S"])

        def ms(ylist, alist, fin):
            a_t = torch.stack(alist, 0)
            suffix = torch.flip(torch.cumsum(torch.flip(a_t, [0]), 0), [0])
            return torch.mean(torch.mean((torch.stack(ylist, 0) + suffix - fin[None, :]) ** 2, dim=-1), dim=-1)
        lh = ms(hys, ahs, model.g(st["hS"]))
        li = ms(ys, as_, model.g(st["S"]))
    else:  # SumLocal :328-364, SumLocalReg :469-505
        lh = li = 0.0
        oh, o = nets(st)
        hYp, Yp = oh[:, 0], o[:, 0]
        for i in range(N):
            if reg:
                a_h, a = -dt * model.f(st["hS"]), -dt * model.f(st["S"])
            else:
                a_h, a = increments(st, oh, o, i, 1)
            tr["hY"].append(hYp); tr["Y"].append(Yp)
            st = model.one_step(st, dW0[i], dW[i], dN[i], hYp, Yp)
            tr["hS"].append(st["hS"]); tr["S"].append(st["S"])
            if i == N - 1:
                hYn, Yn = model.g(st["hS"]), model.g(st["S"])
            else:
                oh, o = nets(st)
                hYn, Yn = oh[:, 0], o[:, 0]
            # reference: (hYNext - hYPrev + toAdd_hat)^2 with toAdd_hat = -a_h   (SumLocal)
            #            (hYPrev - hYNext + toAdd_hat)^2 with toAdd_hat = +a_h   (SumLocalReg) - same square
            lh = lh + torch.mean((hYn - hYp - a_h) ** 2)
            li = li + torch.mean((Yn - Yp - a) ** 2)
            hYp, Yp = hYn, Yn
        tr["hY"].append(hYp); tr["Y"].append(Yp)
    if aux is not None:
        for k, v in tr.items():
            aux[k] = torch.stack(v, 0).detach()
    return lh, li


def sample_mfg_noise(model: MFGOracle, B: int, gen: torch.Generator):
    """Fresh noise drawn the way the reference does: dN_i ~ Poisson(lam(hQ_i) dt) along the exogenous
    hQ path (hQ does not depend on the networks, MFGModel.py:70)."""
    N, sq = model.N, np.float32(np.sqrt(model.dt))
    dW0 = sq * torch.randn(N, B, generator=gen, dtype=model.dtype)
    dW = sq * torch.randn(N, B, generator=gen, dtype=model.dtype)
    hQ = float(model.QAver[0]) * torch.ones(B, dtype=model.dtype)
    dNs = []
    for i in range(N):
        rate = model.lam_dt({"hQ": hQ})
        dNs.append(torch.poisson(rate.clamp_min(0), generator=gen))
        hQ = hQ + model.coeffOU * (float(model.QAver[i + 1]) - hQ) * model.dt + model.sig0 * dW0[i]
    return {"dW0": dW0, "dW": dW, "dN": torch.stack(dNs, 0)}

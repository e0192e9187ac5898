"""Drop-in for coupledMFG/MFGModel.py: the smart-grid mean-field-game model object (MFGModel.py:4-106).

Five states per path: hQ, Q (OU consumption around the load curve QAver, common noise dW0 and idiosyncratic dW),
R (time since the last Cox jump), hS, S (controlled storage of the projected / individual player).  The object is a
parameter holder for the fused kernels (csrc/mfg_kernels.cu); the stateful per-step API of the reference
(`init`, `dN`, `oneStepFrom`, controls, `f`, `g`, state getters) is kept on host tensors for drop-in use and is not on
the training path.  `QAver` is read from the instance (the reference's bare global at :67-68 is a bug).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib as L
from ..runtime import Context, NativeSolver


class ModelCoupledFBSDE:
    def __init__(self, T, QAver, R0, jumpFactor, alpha, beta, coeffOU, A, K, pi, p0, p1, f0, f1, theta, C, S0, h1, h2, sig0,
                 sig, alphaTarget, jumpModel, coeffEqui):
        self.T, self.QAver, self.R0, self.jumpFactor = T, np.asarray(QAver, dtype=np.float64), R0, jumpFactor
        self.alpha, self.beta, self.coeffOU = alpha, beta, coeffOU
        self.A, self.K, self.pi, self.p0, self.p1, self.f0, self.f1, self.theta = A, K, pi, p0, p1, f0, f1, theta
        self.C, self.S0, self.h1, self.h2, self.sig0, self.sig = C, S0, h1, h2, sig0, sig
        self.alphaTarget, self.jumpModel, self.coeffEqui = alphaTarget, jumpModel, coeffEqui
        self.N = len(self.QAver) - 1
        self.dt = T / self.N
        self._gen = torch.Generator().manual_seed(0)

    # ---- native side ---------------------------------------------------------------------------------------
    def c_params(self):
        q = (C.c_double * len(self.QAver))(*self.QAver.tolist())
        p = L.MFGParams(self.T, self.R0, self.jumpFactor, self.alpha, self.beta, self.coeffOU, self.A, self.K, self.pi, self.p0,
                        self.p1, self.f0, self.f1, self.theta, self.C, self.S0, self.h1, self.h2, self.sig0, self.sig,
                        self.alphaTarget, self.coeffEqui, int(self.jumpModel == 'stochastic'), len(self.QAver), q)
        p._keepalive = q
        return p

    def make_solver(self, scheme, nets, n_y0, ctx=None, **kw) -> NativeSolver:
        s = NativeSolver(ctx or Context.default(), L.MODEL_MFG, scheme, nets, n_y0, 0, mfg=self.c_params(), **kw)
        s.N, s.d = self.N, 1
        return s

    # ---- reference per-step API (host tensors) ---------------------------------------------------------------
    def mean_hq(self, i: int) -> float:
        if i == 0:
            return float(self.QAver[0])
        k, dt, j = self.coeffOU, self.dt, np.arange(i)
        return float(math.exp(-k * i * dt) * self.QAver[0] + k * np.sum(self.QAver[:i] * np.exp(k * (j - i) * dt) * dt))

    def init(self, batchSize):
        self.batchSize = batchSize
        one = torch.ones(batchSize, dtype=torch.float32)
        self.hQ, self.Q = float(self.QAver[0]) * one, float(self.QAver[0]) * one
        self.R, self.hS, self.S = self.R0 * one, self.S0 * one, self.S0 * one
        self.meanhQ, self.iStep = float(self.QAver[0]), 0

    def dN(self):
        if self.jumpModel == 'stochastic':
            self.lam = self.beta * (torch.exp(self.alpha * self.hQ) - 1)
        else:
            self.lam = self.jumpFactor * torch.ones(self.batchSize)
        rate = self.lam * self.dt
        return torch.poisson(rate.clamp_min(0.0), generator=self._gen), rate

    def calphaTarget(self):
        if self.jumpModel == 'stochastic':
            return self.alphaTarget * self.meanhQ
        return self.alphaTarget * torch.ones(self.batchSize)

    def _below(self):
        return (self.R <= self.theta).to(torch.float32)

    def calpha_hat(self, hY):
        ce, ind = self.coeffEqui, self._below()
        kTheta = self.A + (1 - self.pi) * ce * self.p1 + self.K + ce * self.f1 * ind
        inner = (self.p0 + self.pi * self.p1 * self.hQ + ((1 - self.pi) * ce * self.p1 + self.K) * self.hQ + hY
                 + (self.f0 + ce * self.f1 * (self.hQ - self.meanhQ - self.calphaTarget())) * ind)
        return -inner / kTheta

    def calpha(self, hY, Y):
        ce, ind, ah = self.coeffEqui, self._below(), self.calpha_hat(hY)
        inner = (self.K * self.Q + self.p0 + self.pi * self.p1 * self.hQ + (1 - self.pi) * ce * self.p1 * (self.hQ + ah) + Y
                 + (self.f0 + ce * self.f1 * (self.hQ - self.meanhQ + ah - self.calphaTarget())) * ind)
        return -inner / (self.A + self.K)

    def oneStepFrom(self, dW0, dW, dN, hY, Y):
        ah, al = self.calpha_hat(hY), self.calpha(hY, Y)     # controls see the states of the current step
        self.iStep += 1
        self.hS = self.hS + ah * self.dt
        self.S = self.S + al * self.dt
        self.R = self.R + self.dt - torch.where(dN > 0, self.R, torch.zeros_like(self.R))
        self.meanhQ = self.mean_hq(self.iStep)
        q = float(self.QAver[self.iStep])
        self.hQ = self.hQ + self.coeffOU * (q - self.hQ) * self.dt + self.sig0 * dW0
        self.Q = self.Q + self.coeffOU * (q - self.Q) * self.dt + self.sig0 * dW0 + self.sig * dW

    def f(self, U):
        return U * self.C

    def g(self, X):
        return self.h1 + self.h2 * X

    def getProjectedStates(self):
        return self.iStep * self.dt, self.hQ, self.hS, self.R

    def getAllStates(self):
        return self.iStep * self.dt, self.Q, self.S, self.hQ, self.hS, self.R

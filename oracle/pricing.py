"""Oracle for the coupled-pricing half of the reference (Merton jump-diffusion, Variance Gamma).

Restates, with injectable noise and torch autograd standing in for tf.GradientTape:
  * coupledPricing/pricingModels.py:10-69   (MertonJumpModel)  -> MertonOracle
  * coupledPricing/pricingModels.py:130-199 (VGmodel, FFT pricer) -> VGOracle
  * coupledPricing/SolversJumpDiff.py       (7 loss graphs)    -> pricing_loss(..., model=MertonOracle)
  * coupledPricing/SolversPureJump.py       (7 loss graphs)    -> pricing_loss(..., model=VGOracle)

State layout: X is [B, d]; the reference is d = 1 ([B]).  The d > 1 Merton model is the
SURVEY.md section 7.4 extension (independent assets, geometric-basket payoff whose closed form
is again a 1-D Merton series); at d = 1 every expression below reduces to the reference's.

Noise is an INPUT (`noise` dict): dW [N,B,d] (Merton only; already sqrt(dt)*N(0,1)),
J [N,B,d] (jump increment of each path), JMC [N,M,d] (the M compensator samples shared by
the batch).  Index i of each tensor is consumed by time step i.  For the SumLocal schemes the
reference draws J_0 before the loop and J_{i+1} at the end of iteration i
(SolversJumpDiff.py:239-240,258-259); the mapping to the same canonical index is 1:1 and the
(N+1)-th draw is never used.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .nets import ParamLayout, mlp_forward

PRICING_SCHEMES = ("Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2", "SumLocalReg", "MultiStepReg")


def _ncdf(x: torch.Tensor) -> torch.Tensor:
    # tfp Normal(0,1).cdf(x) = 0.5 * erfc(-x / sqrt(2))   (TF semantics)
    return 0.5 * torch.erfc(-x * (1.0 / math.sqrt(2.0)))


class MertonOracle:
    """pricingModels.py:10-69.  `aLin` is the slope of the only coupling family the reference
    uses, func(x) = aLin*|x| (mainMerton.py:60-61)."""

    kind = "merton"
    has_brownian = True

    def __init__(self, T, N, r, muJ, sigmaJ, sigma, lam, K, x0, aLin, limit, d: int = 1, dtype=torch.float32):
        self.T, self.N, self.r, self.muJ, self.sigJ, self.sig, self.lam = T, N, r, muJ, sigmaJ, sigma, lam
        self.K, self.x0, self.aLin, self.limit, self.d, self.dtype = K, x0, aLin, limit, d, dtype
        self.dt = T / N
        # geometric-basket equivalents; identical to the inputs when d == 1
        self.sigA = sigma / math.sqrt(d)
        self.lamA = lam * d
        self.muJA = muJ / d
        self.sigJA = sigmaJ / d
        kap = math.exp(muJ + 0.5 * sigmaJ ** 2) - 1.0
        kapA = math.exp(self.muJA + 0.5 * self.sigJA ** 2) - 1.0
        self.qA = 0.0 if d == 1 else (0.5 * sigma ** 2 + lam * kap - 0.5 * self.sigA ** 2 - self.lamA * kapA)

    # pricingModels.py:27-29
    def init(self, B):
        return self.x0 * torch.ones(B, self.d, dtype=self.dtype)

    def drift(self):
        t = lambda v: torch.tensor(v, dtype=self.dtype)
        # (r - sig^2/2 - lam*(exp(muJ + sigJ^2/2) - 1)) : pricingModels.py:54
        return self.r - 0.5 * self.sig * self.sig - self.lam * (torch.exp(t(self.muJ + self.sigJ * self.sigJ * 0.5)) - 1)

    def basket(self, X):
        """Scalar underlying of payoff and closed form: X itself (d=1) or the geometric mean."""
        if self.d == 1:
            return X[:, 0]
        return torch.exp(torch.log(X).mean(dim=1))

    # pricingModels.py:33-49 (BS + A); sigA/lamA/muJA/sigJA/qA == reference parameters when d == 1
    def A(self, iStep, X):
        if iStep >= self.N:
            return self.g(X)
        dt_ = self.dtype
        tau = self.T - iStep * self.dt
        G = self.basket(X)
        if self.d > 1:
            G = G * math.exp(-self.qA * tau)
        I = torch.arange(self.limit, dtype=dt_)
        e = torch.exp(torch.tensor(self.muJA + self.sigJA * self.sigJA * 0.5, dtype=dt_))
        rBS = (self.r - self.lamA * (e - 1) + I * (self.muJA + 0.5 * self.sigJA * self.sigJA) / tau)[None, :]
        sigBS = torch.sqrt(self.sigA ** 2 + I * (self.sigJA ** 2) / tau)[None, :]
        Xt = G[:, None]
        sq = torch.sqrt(torch.tensor(tau, dtype=dt_))
        d1 = (torch.log(Xt / self.K) + (rBS + sigBS ** 2 / 2) * tau) / (sigBS * sq)
        d2 = (torch.log(Xt / self.K) + (rBS - sigBS ** 2 / 2) * tau) / (sigBS * sq)
        BS = Xt * _ncdf(d1) - self.K * torch.exp(-rBS * tau) * _ncdf(d2)
        lam2 = self.lamA * torch.exp(torch.tensor(self.muJA + 0.5 * self.sigJA ** 2, dtype=dt_))
        if self.d == 1:
            coef = (torch.exp(-lam2 * tau) * ((lam2 * tau) ** I) / torch.exp(torch.lgamma(I + 1)))[None, :]
        else:
            # d > 1 needs limit ~ 100 (lamA*T = d*lam*T): the reference's power/factorial quotient overflows
            # fp32 past n = 34, so the same Poisson weight is formed in the log domain in float64.
            l2 = self.lamA * math.exp(self.muJA + 0.5 * self.sigJA ** 2) * tau
            I64 = torch.arange(self.limit, dtype=torch.float64)
            coef = torch.exp(-l2 + I64 * math.log(l2) - torch.lgamma(I64 + 1)).to(dt_)[None, :]
        return (coef * BS).sum(dim=1)

    # pricingModels.py:53-54
    def one_step(self, iStep, X, dW, J, Y):
        coupling = self.aLin * torch.abs(Y - self.A(iStep, X)) * self.dt
        return X * torch.exp(self.drift() * self.dt + self.sig * dW + J) + coupling[:, None]

    def f(self, Y):  # :64-65
        return -self.r * Y

    def g(self, X):  # :68-69
        return torch.clamp_min(self.basket(X) - self.K, 0.0)

    # pricingModels.py:57-61 (sampler; only used by the CPU-baseline timing leg)
    def jumps(self, n, gen: torch.Generator):
        dN = torch.poisson(torch.full((n, self.d), self.lam * self.dt, dtype=self.dtype), generator=gen)
        return dN * self.muJ + self.sigJ * torch.sqrt(dN) * torch.randn(n, self.d, generator=gen, dtype=self.dtype)


class VGOracle:
    """pricingModels.py:130-199 (VGmodel, FFT/Lewis pricer + global cubic spline).

    The per-step spline of `A` depends only on iStep; it is built once in float64
    (numpy ifft + scipy interp1d(kind='cubic'), exactly the reference's calls at :156-175) and, as in
    the reference (tf.numpy_function), its value is a constant for the gradient."""

    kind = "vg"
    has_brownian = False
    d = 1

    def __init__(self, T, N, r, theta, kappa, sigmaJ, K, x0, aLin, dtype=torch.float32):
        self.T, self.N, self.r, self.theta, self.kappa, self.sigJ = T, N, r, theta, kappa, sigmaJ
        self.K, self.x0, self.aLin, self.dtype = K, x0, aLin, dtype
        self.dt = T / N
        self.correction = -math.log(1 - theta * kappa - kappa / 2 * sigmaJ ** 2) / kappa  # :141
        self._splines: Dict[int, object] = {}

    def init(self, B):
        return self.x0 * torch.ones(B, 1, dtype=self.dtype)

    def basket(self, X):
        return X[:, 0]

    @staticmethod
    def fft_grid():
        fftN, Bq = 2 ** 15, 500
        du = Bq / fftN
        rng = np.arange(fftN)
        lm = 2 * np.pi / Bq
        b = fftN * lm / 2
        ku = -b + lm * rng
        return fftN, du, rng, lm, b, ku

    def integral_table(self, iStep) -> np.ndarray:
        """float64 restatement of :152-167: the 2^15 Simpson-weighted Lewis integrand and its ifft."""
        fftN, du, rng, lm, b, ku = self.fft_grid()
        u = rng * du
        tau = self.T - iStep * self.dt
        weight = 3 + (-1.0) ** (rng + 1)
        weight[0], weight[fftN - 1] = 1, 1
        uc = u - 0.5j
        phi = np.exp(tau * (1j * (self.r - self.correction) * uc
                            - np.log(1 - 1j * self.theta * self.kappa * uc + 0.5 * self.kappa * self.sigJ ** 2 * uc * uc) / self.kappa))
        integrand = np.exp(-1j * b * rng * du) * phi / (u ** 2 + 0.25) * weight * du / 3
        return np.real(np.fft.ifft(integrand) * fftN)

    def spline(self, iStep):
        if iStep not in self._splines:
            from scipy.interpolate import interp1d
            _, _, _, _, _, ku = self.fft_grid()
            self._splines[iStep] = interp1d(ku, self.integral_table(iStep), kind="cubic")
        return self._splines[iStep]

    def A(self, iStep, X):
        x = X[:, 0]
        k = torch.log(x / self.K)
        s = torch.as_tensor(self.spline(iStep)(k.detach().double().numpy()), dtype=self.dtype)  # constant for autograd
        tau = self.T - iStep * self.dt
        return x - torch.sqrt(x * self.K) * math.exp(-self.r * tau) / math.pi * s

    def one_step(self, iStep, X, dW, J, Y):  # :184-185
        coupling = self.aLin * torch.abs(Y - self.A(iStep, X)) * self.dt
        return X * torch.exp((self.r - self.correction) * self.dt + J) + coupling[:, None]

    def f(self, Y):
        return -self.r * Y

    def g(self, X):
        return torch.clamp_min(X[:, 0] - self.K, 0.0)

    def jumps(self, n, gen: torch.Generator):  # :188-191 ; tf.random.gamma(alpha=dt/kappa, beta(rate)=1/kappa)
        gauss = torch.randn(n, 1, generator=gen, dtype=self.dtype)
        conc = torch.full((n, 1), self.dt / self.kappa, dtype=torch.float64)
        gam = torch._standard_gamma(conc, generator=gen).to(self.dtype) * self.kappa
        return self.theta * gam + self.sigJ * torch.sqrt(gam) * gauss


def _feat_inputs(model, scheme, t, X, J):
    """Input rows of the jump-term network for jump sample(s) J (broadcast against X)."""
    Xb, Jb = torch.broadcast_tensors(X, J)
    tt = t * torch.ones_like(Xb[..., :1])
    one_net = scheme.endswith("1")
    if model.kind == "merton":
        if scheme == "Global":
            return torch.cat([tt, Xb, Jb], dim=-1)                   # (i, X, J)        SolversJumpDiff.py:37-39
        if one_net:
            return torch.cat([tt, Xb * torch.exp(Jb)], dim=-1)       # (i, X*e^J)       :99-100
        return torch.cat([tt, Xb, torch.exp(Jb)], dim=-1)            # (i, X, e^J)      :173-175
    if one_net:
        return torch.cat([tt, Xb + Xb * Jb], dim=-1)                 # (i, X + X*J)     SolversPureJump.py:95-96
    return torch.cat([tt, Xb, Xb * Jb], dim=-1)                      # (i, X, X*J)      SolversPureJump.py:34-36


def pricing_loss(model, scheme: str, layout: ParamLayout, theta: torch.Tensor, noise: Dict[str, torch.Tensor],
                 B: int, stale_time: bool = True, aux: Optional[dict] = None) -> torch.Tensor:
    """One evaluation of the reference's `optimizeBSDE` / `regressOptim` loss graph.

    Net roles in `layout`: net 0 = UZ / U network, net 1 = Gam network (two-net schemes).
    `theta[layout.y0_offset]` is the trainable Y0 of the Global scheme.
    If `aux` is a dict it receives Y/Z/X trajectories for parity dumps.
    """
    assert scheme in PRICING_SCHEMES
    N, dt, d = model.N, model.dt, model.d
    merton = model.kind == "merton"
    reg = scheme in ("SumLocalReg", "MultiStepReg")
    one_net = scheme.endswith("1")
    gam_net = 0 if one_net else 1
    dW, J, JMC = noise.get("dW"), noise["J"], noise.get("JMC")
    X = model.init(B)
    traj_X, traj_Y, traj_Z = [X], [], []

    def U(t, Xc):
        return mlp_forward(theta, layout, 0, torch.cat([t * torch.ones_like(Xc[:, :1]), Xc], dim=-1))

    def jump_terms(t, Xc, i):
        gam = mlp_forward(theta, layout, gam_net, _feat_inputs(model, scheme, t, Xc, J[i]))[:, 0]
        rows = _feat_inputs(model, scheme, t, Xc[None, :, :], JMC[i][:, None, :])      # [M,B,nin]
        comp = mlp_forward(theta, layout, gam_net, rows)[..., 0].mean(dim=0)
        return gam, comp

    def zdw(Z, i):
        return (Z * dW[i]).sum(dim=1) if merton else 0.0

    if scheme == "Global":
        # SolversJumpDiff.py:22-44 / SolversPureJump.py:22-41
        Y = theta[layout.y0_offset] * torch.ones(B, dtype=theta.dtype)
        for i in range(N):
            t = float(i)
            Z = U(t, X) if merton else None                      # Merton: UZ net outputs Z[d]
            gam, comp = jump_terms(t, X, i)
            traj_Y.append(Y)
            if merton:
                traj_Z.append(Z)
            Y = Y - dt * model.f(Y) + zdw(Z, i) + gam - comp
            X = model.one_step(i, X, dW[i] if merton else None, J[i], Y)   # NEW Y (fact 6 / note ii)
            traj_X.append(X)
        traj_Y.append(Y)
        loss = torch.mean((Y - model.g(X)) ** 2)
    elif scheme in ("MultiStep1", "MultiStep2", "MultiStepReg"):
        # SolversJumpDiff.py:86-115,162-190,461-481 ; SolversPureJump.py same line ranges
        ys, adds = [], []
        for i in range(N):
            t = float(i)
            out = U(t, X)
            Y = out[:, 0]
            if reg:
                add = -dt * model.f(Y)
            else:
                Z = out[:, 1:1 + d] if merton else None
                gam, comp = jump_terms(t, X, i)
                add = -dt * model.f(Y) + zdw(Z, i) + gam - comp
                if merton:
                    traj_Z.append(Z)
            ys.append(Y)
            adds.append(add)
            traj_Y.append(Y)
            X = model.one_step(i, X, dW[i] if merton else None, J[i], Y)
            traj_X.append(X)
        yfin = model.g(X)
        adds_t = torch.stack(adds, 0)
        suffix = torch.flip(torch.cumsum(torch.flip(adds_t, [0]), 0), [0])     # sum_{j>=k} toAdd_j
        F = torch.stack(ys, 0) + suffix
        loss = torch.mean(torch.mean((F - yfin[None, :]) ** 2, dim=-1), dim=-1)
    else:
        # SumLocal1 / SumLocal2 / SumLocalReg : SolversJumpDiff.py:236-269,315-347,391-415
        def tfeat(k):  # time feature fed with state X_k (fact 8: stale after the first step)
            return float(0 if k == 0 else (k - 1 if stale_time else k))
        loss = 0.0
        out = U(tfeat(0), X)
        Yp = out[:, 0]
        for i in range(N):
            if reg:
                add = dt * model.f(Yp)
            else:
                Z = out[:, 1:1 + d] if merton else None
                gam, comp = jump_terms(tfeat(i), X, i)
                add = dt * model.f(Yp) - zdw(Z, i) - gam + comp
                if merton:
                    traj_Z.append(Z)
            traj_Y.append(Yp)
            X = model.one_step(i, X, dW[i] if merton else None, J[i], Yp)
            traj_X.append(X)
            if i == N - 1:
                Yn = model.g(X)
            else:
                out = U(tfeat(i + 1), X)
                Yn = out[:, 0]
            loss = loss + torch.mean((Yn - Yp + add) ** 2)
            Yp = Yn
        traj_Y.append(Yp)
    if aux is not None:
        aux["X"] = torch.stack(traj_X, 0).detach()
        aux["Y"] = torch.stack(traj_Y, 0).detach()
        if traj_Z:
            aux["Z"] = torch.stack(traj_Z, 0).detach()
    return loss


def sample_pricing_noise(model, scheme, B, M, gen: torch.Generator):
    """Fresh noise for one iteration, drawn the way the reference does (used by the CPU baseline)."""
    N, d = model.N, model.d
    noise = {}
    if model.has_brownian:
        noise["dW"] = np.float32(np.sqrt(model.dt)) * torch.randn(N, B, d, generator=gen, dtype=model.dtype)
    noise["J"] = torch.stack([model.jumps(B, gen) for _ in range(N)], 0)
    if scheme not in ("SumLocalReg", "MultiStepReg"):
        noise["JMC"] = torch.stack([model.jumps(M, gen) for _ in range(N)], 0)
    return noise

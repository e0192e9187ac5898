"""Drop-in for coupledPricing/pricingModels.py: the Merton jump-diffusion and Variance-Gamma model objects.

The objects keep the reference's constructor signatures and attributes (pricingModels.py:11-24, :131-143).  They are
parameter holders for the fused sm_100a kernels; the per-step methods the reference's solvers call in their Python
loop (`A`, `jumps`, `oneStepFrom`, `f`, `g`) are kept for drop-in use and evaluate on the GPU through the C-ABI
(`fbsdej_solver_price`, `fbsdej_solver_simulate`) - the training path itself never calls them.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib as L
from ..runtime import Context, NativeSolver, NetSpec


class AbsCoupling:
    """func(x) = aLin * |x| - the only coupling family the reference uses (mainMerton.py:60-61, mainVG.py:57-58)."""

    def __init__(self, aLin: float):
        self.aLin = float(aLin)

    def __call__(self, x):
        return self.aLin * abs(x)


def coupling_slope(func) -> float:
    """Recognise func = a*|x| (a user callable cannot run inside a kernel); raise for anything else."""
    if isinstance(func, AbsCoupling):
        return func.aLin
    if isinstance(func, (int, float)):
        return float(func)
    probe = torch.tensor([-2.0, -0.5, 0.0, 0.5, 2.0], dtype=torch.float64)
    try:
        val = torch.as_tensor(func(probe), dtype=torch.float64)
    except Exception as e:  # e.g. a callable written against the tensorflow API
        raise TypeError("coupling `func` must be AbsCoupling(aLin) or a callable on torch tensors of the form "
                        "a*abs(x)") from e
    a = float(val[-1]) / 2.0
    if not torch.allclose(val, a * probe.abs(), rtol=1e-9, atol=1e-12):
        raise ValueError("only couplings of the form func(x) = a*|x| are compiled into the sm_100a kernels")
    return a


def _as_cpu(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach().to("cpu", torch.float32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32))


class _PricingModel:
    kind = -1
    d = 1

    def _pricer(self) -> NativeSolver:
        if getattr(self, "_native", None) is None:
            self._native = self.make_solver(L.SUMLOCALREG, [NetSpec(1 + self.d, 21, 1), NetSpec(1 + 2 * self.d, 21, 1)], 0, 0)
            self._draw = 0
        return self._native

    def init(self, batchSize):
        self.batchSize = batchSize
        shape = (batchSize,) if self.d == 1 else (batchSize, self.d)
        return self.x0 * torch.ones(shape, dtype=torch.float32)

    def A(self, iStep, X):
        x = _as_cpu(X)
        planes = x.reshape(1, -1) if self.d == 1 else x.reshape(-1, self.d).t().contiguous()
        return torch.from_numpy(self._pricer().price(int(iStep), planes.numpy()))

    def jumps(self, batchSize):
        """One draw of the jump increment for `batchSize` paths (step 0 of a freshly simulated block)."""
        s = self._pricer()
        self._draw += 1
        s.simulate(0x5EED, self._draw, int(batchSize))
        _, ptr_J, *_ = s.get_noise()
        n = int(batchSize) * self.d
        host = np.empty(n, dtype=np.float32)
        L.check(L.lib.fbsdej_memcpy_d2h(s.ctx.handle, host.ctypes.data, ptr_J, n * 4))
        out = torch.from_numpy(host)
        return out if self.d == 1 else out.reshape(self.d, -1).t().contiguous()

    def f(self, Y):
        return -self.r * Y

    def g(self, X):
        x = _as_cpu(X)
        basket = x if self.d == 1 else torch.exp(torch.log(x).mean(dim=-1))
        return torch.clamp_min(basket - self.K, 0.0)


class MertonJumpModel(_PricingModel):
    """MertonJumpModel(T, N, r, muJ, sigmaJ, sigma, lam, K, x0, func, limit)  (pricingModels.py:10-69).

    `d` (keyword, default 1 = the reference) selects the d-asset extension of SURVEY 7.4: independent assets with the
    same parameters and the geometric-basket payoff, whose closed form is again a Merton series."""

    kind = L.MODEL_MERTON

    def __init__(self, T, N, r, muJ, sigmaJ, sigma, lam, K, x0, func, limit, d: int = 1, price_table=None):
        self.T, self.r, self.sig, self.muJ, self.sigJ, self.lam = T, r, sigma, muJ, sigmaJ, lam
        self.K, self.N, self.dt, self.x0, self.func, self.limit, self.d = K, int(N), T / N, x0, func, int(limit), int(d)
        self.aLin = coupling_slope(func)
        # The series of pricingModels.py:40-49 (`limit` terms, two normal CDFs each) is evaluated through the library's per-step
        # cubic-Hermite table in log-moneyness (absolute error < 5e-8; exact series outside the grid) unless price_table=False,
        # which sums the terms at every path-step the way the reference does.
        self.price_table = True if price_table is None else bool(price_table)

    def c_params(self) -> L.MertonParams:
        return L.MertonParams(self.T, self.r, self.muJ, self.sigJ, self.sig, self.lam, self.K, self.x0, self.aLin, self.N,
                              self.limit, self.d)

    def make_solver(self, scheme, nets, n_y0, M, ctx=None, **kw) -> NativeSolver:
        kw.setdefault("price_table", self.price_table)
        s = NativeSolver(ctx or Context.default(), L.MODEL_MERTON, scheme, nets, n_y0, M, merton=self.c_params(), **kw)
        s.N, s.d = self.N, self.d
        return s

    def BS(self, iStep, X, rbs, sigbs):
        """Black-Scholes call (pricingModels.py:33-37), host float32."""
        x, rbs, sigbs = _as_cpu(X), _as_cpu(rbs), _as_cpu(sigbs)
        tau = self.T - iStep * self.dt
        nrm = torch.distributions.Normal(0.0, 1.0)
        d1 = (torch.log(x / self.K) + (rbs + sigbs ** 2 / 2) * tau) / (sigbs * math.sqrt(tau))
        d2 = d1 - sigbs * math.sqrt(tau)
        return x * nrm.cdf(d1) - self.K * torch.exp(-rbs * tau) * nrm.cdf(d2)

    def oneStepFrom(self, iStep, X, dW, gaussJ, Y):
        x, y = _as_cpu(X), _as_cpu(Y)
        mu = self.r - 0.5 * self.sig ** 2 - self.lam * (math.exp(self.muJ + 0.5 * self.sigJ ** 2) - 1.0)
        coup = self.aLin * torch.abs(y - self.A(iStep, x)) * self.dt
        if self.d > 1:
            coup = coup[:, None]
        return x * torch.exp(mu * self.dt + self.sig * _as_cpu(dW) + _as_cpu(gaussJ)) + coup


class VGmodel(_PricingModel):
    """VGmodel(T, N, r, theta, kappa, sigmaJ, K, x0, func)  (pricingModels.py:130-199, Lewis-FFT pricer)."""

    kind = L.MODEL_VG

    def __init__(self, T, N, r, theta, kappa, sigmaJ, K, x0, func):
        self.T, self.r, self.sigJ, self.theta, self.kappa = T, r, sigmaJ, theta, kappa
        self.K, self.N, self.dt, self.x0, self.func = K, int(N), T / N, x0, func
        self.correction = -math.log(1 - theta * kappa - kappa / 2 * sigmaJ ** 2) / kappa
        self.aLin = coupling_slope(func)

    def c_params(self) -> L.VGParams:
        return L.VGParams(self.T, self.r, self.theta, self.kappa, self.sigJ, self.K, self.x0, self.aLin, self.N)

    def make_solver(self, scheme, nets, n_y0, M, ctx=None, **kw) -> NativeSolver:
        s = NativeSolver(ctx or Context.default(), L.MODEL_VG, scheme, nets, n_y0, M, vg=self.c_params(), **kw)
        s.N, s.d = self.N, 1
        return s

    def oneStepFrom(self, iStep, X, gaussJ, Y):
        x, y = _as_cpu(X), _as_cpu(Y)
        coup = self.aLin * torch.abs(y - self.A(iStep, x)) * self.dt
        return x * torch.exp((self.r - self.correction) * self.dt + _as_cpu(gaussJ)) + coup


class VGmodelinvfourier(VGmodel):
    """VGmodelinvfourier(T, N, r, theta, kappa, sigmaJ, K, x0, func)  (pricingModels.py:73-126): the same Variance-Gamma model
    with the call price by DIRECT Fourier inversion, `X Q1 - K e^{-r tau} Q2` with `Q_j = 1/2 + (1/pi) int Re(...) du` on a
    1000-point trapezoid over u in [1e-15, 5000] (:99-107).  It is the independent cross-check of the Lewis/FFT table that
    `VGmodel.A` and the kernels use (0.1331406 vs 0.1331402 at the mainVG.py:54 parameters).

    The reference class cannot be used with any solver (its `jumps()` takes no batch size, :115); here it
    inherits the solver plumbing of VGmodel, so only `A` differs: evaluated on the host in float64 exactly as written."""

    def characteristicfunc(self, iStep, u):
        tau = self.T - iStep * self.dt
        return np.exp(tau * (1j * (self.r - self.correction) * u
                             - np.log(1 - 1j * self.theta * self.kappa * u + 0.5 * self.kappa * self.sigJ * self.sigJ * u * u) / self.kappa))

    def A(self, iStep, X):
        x = np.asarray(_as_cpu(X).numpy(), dtype=np.float64).reshape(-1)
        k = np.log(self.K / x)[None, :]                                   # [1, B]
        u = np.linspace(1e-15, 5000.0, 10 ** 3)[:, None]                  # [1000, 1]
        base = np.exp(-1j * u * k) / (1j * u)
        i1 = np.real(base * self.characteristicfunc(iStep, u - 1j) / self.characteristicfunc(iStep, -1.0000000000001j))
        i2 = np.real(base * self.characteristicfunc(iStep, u + 0j))
        trapz = getattr(np, "trapezoid", None) or np.trapz
        Q1 = 0.5 + trapz(i1, u[:, 0], axis=0) / np.pi
        Q2 = 0.5 + trapz(i2, u[:, 0], axis=0) / np.pi
        price = x * Q1 - self.K * np.exp(-self.r * (self.T - iStep * self.dt)) * Q2
        return torch.from_numpy(price.astype(np.float32)).reshape(_as_cpu(X).shape)

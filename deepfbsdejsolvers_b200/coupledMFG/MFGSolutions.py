"""Drop-in for coupledMFG/MFGSolutions.py: replay of trained networks on PRE-DRAWN increments, the cost functional and the
electricity price (the inputs of the price-of-anarchy study, mainMFGPoA.py:113-121, 189-337).

    MFGSolutionsFixedTrajectory(mathModel, kerasModel, method, dW0_arr, dW_arr, dN)
        .simulateAllProcesses(nbSimulations)  -> attributes R, hQ, meanhQ, Q, lam, hS, S, alpha_hat, alpha, alphaTg
        .price(pi, alpha), .objectiveFunction()

`dW0_arr`, `dW_arr`, `dN` are path-major `[nbSimul, N+1]` arrays as in the reference (MFGSolutions.py:23-31); column i
drives step i, the last column is unused by the state recursion.  The replay itself is ONE forward sweep of the fused MFG
kernel on the injected increments (`fbsdej_solver_set_noise` -> `fbsdej_solver_loss` -> `fbsdej_solver_mfg_states`); the
controls `alpha_hat`, `alpha`, the intensity and the cost are closed-form functions of the dumped states
(MFGModel.py:47-54, 76-89), evaluated on the host in float64.

The reference file does not run as shipped (its constructor reads an undefined `savefig`, MFGSolutions.py:10; SURVEY fact
10); the semantics restated here are those of its `simulateAllProcesses` / `objectiveFunction` bodies.
"""
from __future__ import annotations

import numpy as np

from . import MFGSolvers as _S

_SOLVER_OF = {'Global': _S.SolverGlobalFBSDE, 'SumMultiStep': _S.SolverMultiStepFBSDE, 'SumLocal': _S.SolverSumLocalFBSDE,
              'SumLocalReg': _S.SolverGlobalSumLocalReg, 'SumMultiStepReg': _S.SolverGlobalMultiStepReg}


class MFGSolutionsFixedTrajectory:
    def __init__(self, mathModel, kerasModel, method, dW0_arr, dW_arr, dN, ctx=None):
        if method not in _SOLVER_OF:
            raise ValueError(f"unknown method {method!r}; expected one of {sorted(_SOLVER_OF)}")
        self.mathModel, self.kerasModel, self.method = mathModel, kerasModel, method
        self.dW0_arr, self.dW_arr, self.dN = (np.asarray(x, dtype=np.float32) for x in (dW0_arr, dW_arr, dN))
        self.t = np.arange(self.mathModel.N + 1)
        self.dt = self.mathModel.dt
        self.theta = self.mathModel.theta
        self._solver = _SOLVER_OF[method](mathModel, kerasModel, 0.0, 'ON', ctx=ctx)     # carries the trained parameters

    # -- closed-form controls on arrays of states (MFGModel.py:76-89) --------------------------------------------
    def _controls(self, hQ, Q, R, meanhQ, alphaTg, hY, Y):
        m = self.mathModel
        ce, ind = m.coeffEqui, (R <= m.theta).astype(np.float64)
        kTheta = m.A + (1 - m.pi) * ce * m.p1 + m.K + ce * m.f1 * ind
        ah = -(m.p0 + m.pi * m.p1 * hQ + ((1 - m.pi) * ce * m.p1 + m.K) * hQ + hY
               + (m.f0 + ce * m.f1 * (hQ - meanhQ - alphaTg)) * ind) / kTheta
        al = -(m.K * Q + m.p0 + m.pi * m.p1 * hQ + (1 - m.pi) * ce * m.p1 * (hQ + ah) + Y
               + (m.f0 + ce * m.f1 * (hQ - meanhQ + ah - alphaTg)) * ind) / (m.A + m.K)
        return ah, al

    def simulateAllProcesses(self, nbSimulations):
        if nbSimulations > self.dN.shape[0]:
            raise Exception('Shape error, choose a number of simulations lower than the shape dN.')
        m, B = self.mathModel, int(nbSimulations)
        N = m.N
        s = self._solver.build()
        planes = [np.ascontiguousarray(x[:B, :N].T) for x in (self.dW0_arr, self.dW_arr, self.dN)]     # [N, B]
        s.set_noise(B, *planes)
        _, _, ty, _ = s.loss(B, traj=True)                      # (hY, Y) fed to the controls at steps 0 .. N-1
        st = s.mfg_states(B).astype(np.float64)                 # [N+1, 5, B]: hQ, Q, R, hS, S
        self.hQ, self.Q, self.R, self.hS, self.S = (np.ascontiguousarray(st[:, k, :].T) for k in range(5))
        self.meanhQ = np.array([m.mean_hq(i) for i in range(N + 1)])
        if m.jumpModel == 'stochastic':
            self.lam = m.beta * (np.exp(m.alpha * self.hQ) - 1.0)                                   # MFGModel.py:49
            self.alphaTg = m.alphaTarget * np.tile(self.meanhQ[None, :], (B, 1))
        else:
            self.lam = m.jumpFactor * np.ones((B, N + 1))
            self.alphaTg = m.alphaTarget * np.ones((B, N + 1))
        self.lam[:, N] = 0.0                                    # the reference never fills the last column (:84)
        hY, Y = ty[:, 0, :].astype(np.float64).T.copy(), ty[:, 1, :].astype(np.float64).T.copy()      # [B, N+1]
        if self.method != 'Global':
            # the non-Global replays evaluate the networks on the terminal state as well (:97-98); the kernel's dump holds
            # g(S_N) there, so that one column comes from a direct network call
            tN = np.full(B, N * m.dt)
            xh = np.stack([tN, self.hQ[:, N], self.hS[:, N], self.R[:, N]], 1).astype(np.float32)
            xi = np.stack([tN, self.Q[:, N], self.S[:, N], self.hQ[:, N], self.hS[:, N], self.R[:, N]], 1).astype(np.float32)
            hY[:, N], Y[:, N] = s.net_forward(0, xh)[:, 0], s.net_forward(1, xi)[:, 0]
        self.alpha_hat, self.alpha = self._controls(self.hQ, self.Q, self.R, self.meanhQ[None, :], self.alphaTg, hY, Y)
        self.hY, self.Y = hY, Y

    def price(self, pi, alpha):
        m = self.mathModel
        return m.p0 + pi * m.p1 * self.hQ + (1 - pi) * m.p1 * (self.hQ + alpha)

    def objectiveFunction(self):
        m = self.mathModel
        increment = (m.A * 0.5 * self.alpha ** 2 + m.C * 0.5 * self.S ** 2 + m.K * 0.5 * (self.Q + self.alpha) ** 2
                     + (self.Q + self.alpha) * (m.p0 + m.p1 * m.pi * self.hQ + m.p1 * (1 - m.pi) * (self.hQ + self.alpha_hat))
                     + (self.R < m.theta) * (self.Q - self.meanhQ + self.alpha - self.alphaTg)
                     * (m.f0 + m.f1 * (self.hQ - self.meanhQ + self.alpha_hat - self.alphaTg)))
        cost_integral = np.sum(increment * m.dt, axis=1) + m.h1 * self.S[:, -1] + m.h2 * 0.5 * self.S[:, -1] ** 2
        return np.mean(cost_integral), np.std(cost_integral)

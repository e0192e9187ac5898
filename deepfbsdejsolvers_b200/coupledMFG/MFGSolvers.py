"""Drop-in for coupledMFG/MFGSolvers.py: the five smart-grid MFG solver classes.

`Solver*(mathModel, modelKeras, lRate, couplage).train(batchSize, batchSizeVal, num_epoch, num_epochExt)` returns
(listY0_hat, listY0) as in the reference (MFGSolvers.py:116).  couplage 'ON' trains both networks on the summed
objective (:66-73); 'OFF' trains the projected player on its own loss, then the individual player on its own loss with
the SAME optimizer object, i.e. the Adam step counter carries over (:75, :92-115) - reproduced with objective weights
and a parameter mask.  The loss graphs run in csrc/mfg_kernels.cu.
"""
from __future__ import annotations

import time
from typing import List

import numpy as np
import torch

from .. import _lib as L
from ..solver_base import TrainLoop, dist_info


class SolverBase:
    SCHEME = L.GLOBAL
    REG = False

    def __init__(self, mathModel, modelKeras, lRate, couplage, seed: int = 0, ctx=None, tensor_cores=None):
        self.mathModel, self.modelKeras, self.lRate, self.couplage = mathModel, modelKeras, lRate, couplage
        self.seed, self.ctx, self.native = seed, ctx, None
        # tcgen05 kernels (csrc/mfg_tc_kernels.cu; tests/test_tc_gpu.py): None = automatic, on when both networks fit them
        self.tensor_cores = tensor_cores

    # expected output widths (mainMFGComparison.py:119-124)
    def _widths(self):
        if self.SCHEME == L.GLOBAL:
            return 2, 3
        return (1, 1) if self.REG else (3, 4)

    def build(self):
        if self.native is not None:
            return self.native
        hat, ind = self.modelKeras.model_hat, self.modelKeras.model
        hat.build(4)
        ind.build(6)
        wh, wi = self._widths()
        if (hat.ndimOut, ind.ndimOut) != (wh, wi):
            raise ValueError(f"{type(self).__name__}: networks must have ndimOut ({wh}, {wi}), got ({hat.ndimOut}, {ind.ndimOut})")
        n_y0 = 2 if self.SCHEME == L.GLOBAL else 0
        sh, si = hat.spec(), ind.spec()
        tc_ok = sh.H <= 22 and si.H <= 22 and sh.L == 2 and si.L == 2 and sh.activation == si.activation
        if self.tensor_cores and not tc_ok:
            raise ValueError(f"{type(self).__name__}: tensor_cores=True, but the tcgen05 MFG kernels need two hidden layers of width "
                             f"<= 22 and one activation for both networks; got {sh} and {si}")
        use_tc = tc_ok if self.tensor_cores is None else bool(self.tensor_cores)
        self.native = self.mathModel.make_solver(self.SCHEME, [sh, si], n_y0, ctx=self.ctx, tensor_cores=use_tc)
        parts = [hat.params, ind.params]
        if n_y0:
            parts.append(np.array([hat.Y0_hat.numpy(), ind.Y0.numpy()], dtype=np.float32))
        self.native.set_theta(np.concatenate(parts))
        return self.native

    def pull_params(self):
        s, hat, ind = self.native, self.modelKeras.model_hat, self.modelKeras.model
        th = s.get_theta()
        hat.params = th[s.offsets[0]:s.offsets[0] + hat.params.size].copy()
        ind.params = th[s.offsets[1]:s.offsets[1] + ind.params.size].copy()
        if self.SCHEME == L.GLOBAL:
            hat.Y0_hat.assign(th[s.y0_offset])
            ind.Y0.assign(th[s.y0_offset + 1])

    def current_Y0(self):
        s, mm = self.native, self.mathModel
        if self.SCHEME == L.GLOBAL:
            th = s.get_theta()
            return float(th[s.y0_offset]), float(th[s.y0_offset + 1])
        q0 = float(mm.QAver[0])   # init(1) states, MFGSolvers.py:264-265
        yh = s.net_forward(0, np.array([[0.0, q0, mm.S0, mm.R0]], dtype=np.float32))[0, 0]
        yi = s.net_forward(1, np.array([[0.0, q0, mm.S0, q0, mm.S0, mm.R0]], dtype=np.float32))[0, 0]
        return float(yh), float(yi)

    # ---- diagnostics (MFGSolvers.py:118-178, 296-318, 436-459, 581-602, 727-748): forward-only replays of the trained
    # networks on fresh increments; the sweep itself is the forward kernel with its trajectory dump -----------------------
    SEED_DIAG = 0x64696167

    def _replay(self, nbSimul: int, noise=None):
        """noise = (dW0, dW, dN), each [N, nbSimul]: replay on given increments (extension: the reference always draws fresh
        ones); None: fresh device draws."""
        s = self.build()
        if noise is not None:
            s.set_noise(nbSimul, *[np.ascontiguousarray(x, dtype=np.float32) for x in noise])
        else:
            self._ndiag = getattr(self, "_ndiag", 0) + 1
            s.simulate(self.seed ^ self.SEED_DIAG, self._ndiag, nbSimul)
        out, tx, ty, _ = s.loss(nbSimul, traj=True)          # tx: (hS, S) [N+1, 2, B]; ty: (hY, Y) [N+1, 2, B]
        return s, out, tx.astype(np.float64), ty.astype(np.float64)

    def simulateGlobalErr(self, nbSimul, noise=None):
        """(mean cost of the projected player, mean cost of the individual player, terminal mismatch): cost =
        sum_i dt f(S_i) + g(S_N) with f(U) = C U, g(X) = h1 + h2 X (MFGModel.py:92-98)."""
        s, out, tx, ty = self._replay(nbSimul, noise)
        mm = self.mathModel
        N, dt = s.N, float(mm.T) / s.N
        run = dt * float(mm.C) * tx[:N].sum(axis=0)                                   # [2, B]
        gN = float(mm.h1) + float(mm.h2) * tx[N]                                      # [2, B]
        cost = (run + gN).mean(axis=1)
        last = ty[N] if self.SCHEME == L.GLOBAL else ty[N - 1]                       # others compare the LAST net output (:318)
        mismatch = float(((last - gN) ** 2).mean(axis=1).sum())
        return float(cost[0]), float(cost[1]), mismatch

    def followS(self, nbSimul, noise=None):
        """Mean and (population) standard deviation of hS and S at every time step (MFGSolvers.py:148-178)."""
        _, _, tx, _ = self._replay(nbSimul, noise)
        return (list(tx[:, 0].mean(axis=1)), list(tx[:, 0].std(axis=1)), list(tx[:, 1].mean(axis=1)), list(tx[:, 1].std(axis=1)))

    # checkpoint / resume: parameters | Adam m, v, t | Philox iteration; the reference keeps nothing on disk
    def save(self, path: str) -> None:
        self.build().save_checkpoint(path, seed=self.seed, listY0_hat=getattr(self, "listY0_hat", []),
                                     listY0=getattr(self, "listY0", []), lossList=getattr(self, "lossList", []))

    def load(self, path: str) -> None:
        sd = self.build().load_checkpoint(path)
        self.seed = int(sd.get("seed", self.seed))
        self.listY0_hat = [np.float32(x) for x in sd.get("listY0_hat", [])]
        self.listY0 = [np.float32(x) for x in sd.get("listY0", [])]
        self.lossList = [float(x) for x in sd.get("lossList", [])]
        self._resumed = True               # the next train() keeps the restored optimizer state
        self.pull_params()

    def _mask(self, which: str) -> torch.Tensor:
        s = self.native
        m = np.zeros(s.P, dtype=np.float32)
        n0 = self.modelKeras.model_hat.params.size
        if which == "hat":
            m[:n0] = 1
            if self.SCHEME == L.GLOBAL:
                m[s.y0_offset] = 1
        else:
            m[s.offsets[1]:s.y0_offset] = 1
            if self.SCHEME == L.GLOBAL:
                m[s.y0_offset + 1] = 1
        return s.ctx.to_device(m)

    def train(self, batchSize, batchSizeVal, num_epoch, num_epochExt):
        s = self.build()
        loop = TrainLoop(s, self.lRate, self.seed)
        resumed, self._resumed = getattr(self, "_resumed", False), False
        if not resumed:                    # one optimizer object per train() call (MFGSolvers.py:75) unless load() restored one
            s.reset_optimizer()
            self.listY0_hat: List[float] = []
            self.listY0: List[float] = []
            self.lossList: List[float] = []
        _, rank, _ = dist_info()
        draw = 0

        def phase(w_hat, w_ind, mask, idx, label):
            nonlocal draw
            s.set_weights(w_hat, w_ind)
            for iout in range(num_epochExt):
                t0 = time.time()
                loop.steps(batchSize, num_epoch, mask=mask)
                rtime = time.time() - t0
                err = float(loop.validation(batchSizeVal, draw)[idx])
                draw += 1
                yh, yi = self.current_Y0()
                if rank == 0:
                    print(label, err, " took %5.3f s" % rtime, "Y0_hat sofar ", yh, 'Y0 sofar', yi, 'epoch', iout)
                self.lossList.append(err)
                if idx in (0, 1):
                    self.listY0_hat.append(np.float32(yh))
                if idx in (0, 2):
                    self.listY0.append(np.float32(yi))

        if self.couplage == 'ON':
            phase(1.0, 1.0, None, 0, "Error ")
        else:
            phase(1.0, 0.0, self._mask("hat"), 1, "Error hat ")
            phase(0.0, 1.0, self._mask("ind"), 2, " Error")
        s.set_weights(1.0, 1.0)
        self.pull_params()
        return self.listY0_hat, self.listY0


class SolverGlobalFBSDE(SolverBase):
    """MFGSolvers.py:17-116."""
    SCHEME = L.GLOBAL


class SolverMultiStepFBSDE(SolverBase):
    """MFGSolvers.py:180-294."""
    SCHEME = L.MULTISTEP2


class SolverSumLocalFBSDE(SolverBase):
    """MFGSolvers.py:321-434."""
    SCHEME = L.SUMLOCAL2


class SolverGlobalSumLocalReg(SolverBase):
    """MFGSolvers.py:463-579."""
    SCHEME, REG = L.SUMLOCALREG, True


class SolverGlobalMultiStepReg(SolverBase):
    """MFGSolvers.py:608-725."""
    SCHEME, REG = L.MULTISTEPREG, True

"""Drop-in for coupledPricing/SolversPureJump.py: the seven pure-jump (Variance Gamma) solver classes.

No Brownian part, jump feature X*J (two networks) or X + X*J (one network); the trainable Y0 of the Global scheme
lives on the Gam network (SolversPureJump.py:27,67); `train` returns (listY0, durationList) (:72).
"""
from __future__ import annotations

from .. import _lib as L
from ..solver_base import PricingSolverBase


class SolverBase(PricingSolverBase):
    """SolversPureJump.py:6-15."""

    def __init__(self, mathModel, modelKerasU, modelKerasGam, lRate, **kw):
        super().__init__(mathModel, modelKerasU, modelKerasGam, lRate, **kw)
        self.modelKerasU, self.modelKerasGam = modelKerasU, modelKerasGam

    def _result(self):
        return self.listY0, self.durationList


class _OneNet(PricingSolverBase):
    TWO_NET = False

    def __init__(self, mathModel, modelKerasU, lRate, **kw):
        super().__init__(mathModel, modelKerasU, None, lRate, **kw)
        self.modelKerasU = modelKerasU

    def _result(self):
        return self.listY0, self.durationList


class SolverGlobalFBSDE(SolverBase):
    """SolversPureJump.py:17-72."""
    SCHEME, Y0_NET = L.GLOBAL, "Gam"


class SolverMultiStepFBSDE1(_OneNet):
    """SolversPureJump.py:74-141."""
    SCHEME = L.MULTISTEP1


class SolverMultiStepFBSDE2(SolverBase):
    """SolversPureJump.py:143-208."""
    SCHEME = L.MULTISTEP2


class SolverSumLocalFBSDE1(_OneNet):
    """SolversPureJump.py:210-280."""
    SCHEME = L.SUMLOCAL1


class SolverSumLocalFBSDE2(SolverBase):
    """SolversPureJump.py:282-351."""
    SCHEME = L.SUMLOCAL2


class SolverGlobalSumLocalReg(SolverBase):
    """SolversPureJump.py:355-414 (train 1000*batchSize :403, validation 100*batchSizeVal :407)."""
    SCHEME, REG, TRAIN_MULT, VAL_MULT = L.SUMLOCALREG, True, 1000, 100


class SolverGlobalMultiStepReg(SolverBase):
    """SolversPureJump.py:422-482 (train 1000*batchSize :471, validation 100*batchSizeVal :475)."""
    SCHEME, REG, TRAIN_MULT, VAL_MULT = L.MULTISTEPREG, True, 1000, 100

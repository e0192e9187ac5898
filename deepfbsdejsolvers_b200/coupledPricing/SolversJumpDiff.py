"""Drop-in for coupledPricing/SolversJumpDiff.py: the seven jump-diffusion (Merton) solver classes.

Constructor signatures and `.train(batchSize, batchSizeVal, num_epoch, num_epochExt)` follow the reference; the
loss graphs themselves run in the fused sm_100a kernels (csrc/pricing_kernels.cu), selected by SCHEME.
Extra keyword arguments (`M`, `seed`, `ctx`, `stale_time`) expose what the reference hard-codes.
"""
from __future__ import annotations

from .. import _lib as L
from ..solver_base import PricingSolverBase


class SolverGlobalFBSDE(PricingSolverBase):
    """SolversJumpDiff.py:17-73.  Trainable Y0 lives on modelKerasUZ (:27, :69)."""
    SCHEME, TWO_NET, Y0_NET = L.GLOBAL, True, "UZ"

    def __init__(self, mathModel, modelKerasUZ, modelKerasGam, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, modelKerasGam, lRate, **kw)
        self.modelKerasUZ, self.modelKerasGam = modelKerasUZ, modelKerasGam


class SolverMultiStepFBSDE1(PricingSolverBase):
    """SolversJumpDiff.py:75-149 (one network; jump term = net(i, X e^J)[0])."""
    SCHEME, TWO_NET = L.MULTISTEP1, False

    def __init__(self, mathModel, modelKerasUZ, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, None, lRate, **kw)
        self.modelKerasUZ = modelKerasUZ


class SolverMultiStepFBSDE2(PricingSolverBase):
    """SolversJumpDiff.py:151-224."""
    SCHEME, TWO_NET = L.MULTISTEP2, True

    def __init__(self, mathModel, modelKerasUZ, modelKerasGam, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, modelKerasGam, lRate, **kw)
        self.modelKerasUZ, self.modelKerasGam = modelKerasUZ, modelKerasGam


class SolverSumLocalFBSDE1(PricingSolverBase):
    """SolversJumpDiff.py:226-303."""
    SCHEME, TWO_NET = L.SUMLOCAL1, False

    def __init__(self, mathModel, modelKerasUZ, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, None, lRate, **kw)
        self.modelKerasUZ = modelKerasUZ


class SolverSumLocalFBSDE2(PricingSolverBase):
    """SolversJumpDiff.py:305-381."""
    SCHEME, TWO_NET = L.SUMLOCAL2, True

    def __init__(self, mathModel, modelKerasUZ, modelKerasGam, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, modelKerasGam, lRate, **kw)
        self.modelKerasUZ, self.modelKerasGam = modelKerasUZ, modelKerasGam


class SolverGlobalSumLocalReg(PricingSolverBase):
    """SolversJumpDiff.py:385-445: no Z / Gam / compensator; train batch 1000*batchSize (:435), validation
    100*batchSizeVal (:439)."""
    SCHEME, TWO_NET, REG, TRAIN_MULT, VAL_MULT = L.SUMLOCALREG, True, True, 1000, 100

    def __init__(self, mathModel, modelKerasUZ, modelKerasGam, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, modelKerasGam, lRate, **kw)
        self.modelKerasUZ, self.modelKerasGam = modelKerasUZ, modelKerasGam


class SolverGlobalMultiStepReg(PricingSolverBase):
    """SolversJumpDiff.py:453-513: train batch 1000*batchSize (:503), validation batchSizeVal (:507)."""
    SCHEME, TWO_NET, REG, TRAIN_MULT, VAL_MULT = L.MULTISTEPREG, True, True, 1000, 1

    def __init__(self, mathModel, modelKerasUZ, modelKerasGam, lRate, **kw):
        super().__init__(mathModel, modelKerasUZ, modelKerasGam, lRate, **kw)
        self.modelKerasUZ, self.modelKerasGam = modelKerasUZ, modelKerasGam

"""Drop-in for the reference's coupledPricing/ directory (Merton jump-diffusion and Variance-Gamma pricing)."""
from .Networks import Net  # noqa: F401
from .pricingModels import MertonJumpModel, VGmodel, VGmodelinvfourier, AbsCoupling  # noqa: F401
from .SolversJumpDiff import (SolverGlobalFBSDE, SolverMultiStepFBSDE1, SolverMultiStepFBSDE2, SolverSumLocalFBSDE1,  # noqa: F401
                              SolverSumLocalFBSDE2, SolverGlobalSumLocalReg, SolverGlobalMultiStepReg)

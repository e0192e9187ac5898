"""Drop-in for the reference's coupledMFG/ directory (smart-grid mean-field game with Cox-process jumps)."""
from .Networks import Net_hat, Net, kerasModels  # noqa: F401
from .MFGModel import ModelCoupledFBSDE  # noqa: F401
from .MFGSolutions import MFGSolutionsFixedTrajectory  # noqa: F401
from .MFGSolvers import (SolverGlobalFBSDE, SolverMultiStepFBSDE, SolverSumLocalFBSDE, SolverGlobalSumLocalReg,  # noqa: F401
                         SolverGlobalMultiStepReg)

"""deepfbsdejsolvers_b200 - B200 (sm_100a) implementation of the deep-FBSDE-with-jumps training hot path of
ZakariaBensaid/DeepFBSDEJSolvers, behind the reference's Python class API.

    from deepfbsdejsolvers_b200.coupledPricing import MertonJumpModel, Net, SolverGlobalFBSDE, AbsCoupling
    from deepfbsdejsolvers_b200.coupledMFG import ModelCoupledFBSDE, kerasModels, Net_hat, Net, SolverGlobalFBSDE

Importing the package loads libfbsdej.so (C-ABI in include/fbsdej.h) and raises if it has not been built;
there is no CPU fallback.
"""
from ._lib import FbsdejError, LIB_PATH, SYMBOLS  # noqa: F401
from .init import set_seed  # noqa: F401
from .runtime import Context, NativeSolver, NetSpec  # noqa: F401

__all__ = ["Context", "NativeSolver", "NetSpec", "FbsdejError", "set_seed", "LIB_PATH", "SYMBOLS"]

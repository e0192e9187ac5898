"""Drop-in for coupledMFG/Networks.py: `Net_hat`, `Net`, `kerasModels` of the smart-grid mean-field game."""
from __future__ import annotations

import numpy as np
import torch

from .. import init as _init
from ..nets import DenseNet, Scalar

_NO_Y0 = ('SumLocal', 'SumMultiStep', 'SumMultiStepReg', 'SumLocalReg', 'Osterlee')


class _StateNet(DenseNet):
    def __call__(self, inputs):
        """Networks.py:17-21 / :35-39: `inputs` is the state tuple (t, s1, s2, ...) with scalar time."""
        cols = [np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, dtype=np.float32) for x in inputs]
        n = max(c.size for c in cols[1:])
        stacked = np.stack([np.broadcast_to(c.reshape(-1) if c.size > 1 else c.reshape(()), (n,)) for c in cols], axis=-1)
        return super().__call__(stacked)


class Net_hat(_StateNet):
    """Network of the common-noise-projected player: inputs (t, hQ, hS, R); Y0_hat ~ GlorotUniform([])."""

    def __init__(self, method, ndimOut, nbNeurons, activation="tanh"):
        super().__init__(ndimOut, nbNeurons, activation)
        self.name_ = "FeedForwardBSDEProjectedCase"
        if method not in _NO_Y0:
            self.Y0_hat = Scalar(_init.glorot_uniform(()))


class Net(_StateNet):
    """Network of the individual player: inputs (t, Q, S, hQ, hS, R); Y0 ~ GlorotNormal([])."""

    def __init__(self, method, ndimOut, nbNeurons, activation="tanh"):
        super().__init__(ndimOut, nbNeurons, activation)
        self.name_ = "FeedForwardBSDE"
        if method not in _NO_Y0:
            self.Y0 = Scalar(_init.glorot_normal(()))


class kerasModels:
    """Pair container (Networks.py:42-46)."""

    def __init__(self, Net_hat, Net, method, ndimOut_hat, ndimOut, nbNeurons_hat, nbNeurons, activation_hat, activation="tanh"):
        self.model_hat = Net_hat(method, ndimOut_hat, nbNeurons_hat, activation_hat)
        self.model = Net(method, ndimOut, nbNeurons, activation)

"""Drop-in for coupledPricing/Networks.py of the reference: `Net(bY0, ndimOut, nbNeurons, activation)`."""
from __future__ import annotations

from .. import init as _init
from ..nets import DenseNet, Scalar


class Net(DenseNet):
    """Networks.py:6-23.  `.Y0` exists iff bY0 == 1 (GlorotNormal([]) scalar, Networks.py:14-15)."""

    def __init__(self, bY0, ndimOut, nbNeurons, activation="tanh"):
        super().__init__(ndimOut, nbNeurons, activation)
        self.name_ = "FeedForwardANd0"
        if bY0 == 1:
            self.Y0 = Scalar(_init.glorot_normal(()))

"""Host-side network objects shared by coupledPricing.Networks and coupledMFG.Networks.

A network is `nin -> H (act) -> H (act) -> nout` (reference: Dense stacks, coupledPricing/Networks.py:6-23).  Like a
Keras model it is built lazily: the input width is only known when the solver (or the first call) supplies it.
Parameters live on the host as one flat float32 vector in the library layout (per layer W[in][out] row-major, then b);
solvers copy them to the device for training and write the trained values back.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import init as _init
from .runtime import Context, NetSpec


def activation_name(activation) -> str:
    if isinstance(activation, str):
        name = activation
    else:
        name = getattr(activation, "__name__", str(activation))
    name = name.lower()
    for k in ("tanh", "relu"):
        if k in name:
            return k
    raise ValueError(f"activation {activation!r} is not supported (tanh | relu)")


class Scalar:
    """Stand-in for a trainable scalar tf.Variable (`.numpy()` as used at SolversJumpDiff.py:69-70)."""

    def __init__(self, value: float):
        self.value = np.float32(value)

    def numpy(self):
        return np.float32(self.value)

    def assign(self, v):
        self.value = np.float32(v)

    def __float__(self):
        return float(self.value)

    def __repr__(self):
        return f"Scalar({float(self.value):.8g})"


class DenseNet:
    def __init__(self, ndimOut: int, nbNeurons, activation="tanh"):
        self.nbNeurons = [int(h) for h in np.asarray(nbNeurons).reshape(-1)]
        if not 1 <= len(self.nbNeurons) <= 3 or len(set(self.nbNeurons)) != 1:
            raise ValueError("the sm_100a kernels take 1, 2 or 3 EQUAL hidden layers (the reference's nbLayer / nbNeuron; default 2 x 21) "
                             "of width <= 35 for the compensator-free solvers, <= 31 otherwise (three layers: <= 23); the tcgen05 "
                             "kernels cover two layers of width <= 22")
        self.ndimOut = int(ndimOut)
        self.activation = activation_name(activation)
        self.params: Optional[np.ndarray] = None
        self.nin: Optional[int] = None

    @property
    def H(self) -> int:
        return self.nbNeurons[0]

    @property
    def L(self) -> int:
        return len(self.nbNeurons)

    def spec(self) -> NetSpec:
        assert self.nin is not None, "network not built yet"
        return NetSpec(self.nin, self.H, self.ndimOut, self.activation, self.L)

    def build(self, nin: int) -> None:
        """Create the variables (kernels Glorot-normal, biases zero) for input width nin; no-op if already built."""
        if self.params is not None:
            if nin != self.nin:
                raise ValueError(f"network was built for {self.nin} inputs, got {nin}")
            return
        self.nin = int(nin)
        dims = [self.nin] + [self.H] * self.L + [self.ndimOut]
        parts = []
        for a, b in zip(dims[:-1], dims[1:]):
            parts += [_init.glorot_normal((a, b)).reshape(-1), np.zeros(b, dtype=np.float32)]
        self.params = np.concatenate(parts).astype(np.float32)

    def __call__(self, inputs) -> List[torch.Tensor]:
        """Net.call (Networks.py:17-23): stacked inputs [..., nin] -> list of output columns (CPU tensors)."""
        x = inputs.detach().cpu().numpy() if isinstance(inputs, torch.Tensor) else np.asarray(inputs)
        x = np.asarray(x, dtype=np.float32)
        self.build(x.shape[-1])
        ctx = Context.default()
        y = ctx.net_forward(ctx.to_device(self.params), self.spec(), ctx.to_device(x.reshape(-1, x.shape[-1])))
        y = ctx.to_host(y).reshape(*x.shape[:-1], self.ndimOut)
        return [y[..., i] for i in range(self.ndimOut)]

    def layer_arrays(self):
        """[(W, b), ...] views of the flat parameter vector."""
        dims = [self.nin] + [self.H] * self.L + [self.ndimOut]
        out, off = [], 0
        for a, b in zip(dims[:-1], dims[1:]):
            W = self.params[off:off + a * b].reshape(a, b); off += a * b
            bb = self.params[off:off + b]; off += b
            out.append((W, bb))
        return out

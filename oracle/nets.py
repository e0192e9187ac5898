"""Oracle networks: Dense MLPs over a flat fp32 parameter vector.

Follows coupledPricing/Networks.py:6-23 and coupledMFG/Networks.py:6-46 of the
reference: `Dense(H, act)` x L then `Dense(nout)`; Dense is `act(x @ W[in,out] + b)`
(TF semantics); kernel initialiser GlorotNormal (truncated normal, std
sqrt(2/(fan_in+fan_out))/0.87962566), bias zero; optional scalar `Y0`
(GlorotNormal([]) for pricing / MFG `Net`, GlorotUniform([]) for MFG `Net_hat`).

Flat parameter layout (shared with the CUDA library, include/fbsdej.h):
per net, per layer `W[in,out]` row-major then `b[out]`; nets concatenated in
order; `Y0` scalars last.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np
import torch

_TRUNC_STD = 0.87962566103423978


@dataclass
class MLPSpec:
    nin: int
    hidden: Sequence[int]
    nout: int
    activation: str = "tanh"  # 'tanh' | 'relu'

    @property
    def dims(self) -> List[int]:
        return [self.nin, *[int(h) for h in self.hidden], self.nout]

    @property
    def nparams(self) -> int:
        d = self.dims
        return sum(d[i] * d[i + 1] + d[i + 1] for i in range(len(d) - 1))


@dataclass
class ParamLayout:
    """Flat layout of several nets followed by `n_y0` trainable scalars."""

    nets: List[MLPSpec]
    n_y0: int = 0
    offsets: List[int] = field(default_factory=list)

    def __post_init__(self):
        off = 0
        self.offsets = []
        for n in self.nets:
            self.offsets.append(off)
            off += n.nparams
        self.y0_offset = off
        self.total = off + self.n_y0

    def net_slices(self, k: int):
        """[(w_start, w_end, in, out, b_start, b_end), ...] for net k."""
        d = self.nets[k].dims
        off = self.offsets[k]
        out = []
        for i in range(len(d) - 1):
            w0, w1 = off, off + d[i] * d[i + 1]
            b0, b1 = w1, w1 + d[i + 1]
            out.append((w0, w1, d[i], d[i + 1], b0, b1))
            off = b1
        return out


def glorot_normal(shape, rng: np.random.Generator) -> np.ndarray:
    """Keras GlorotNormal: truncated normal (|z| <= 2), std sqrt(2/(fi+fo))/0.8796 (TF semantics)."""
    if len(shape) == 0:
        fi = fo = 1
    elif len(shape) == 1:
        fi = fo = shape[0]
    else:
        fi, fo = shape[0], shape[1]
    std = np.sqrt(2.0 / (fi + fo)) / _TRUNC_STD
    n = int(np.prod(shape)) if len(shape) else 1
    z = rng.standard_normal(n)
    bad = np.abs(z) > 2.0
    while bad.any():
        z[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2.0
    return (std * z).reshape(shape).astype(np.float32)


def glorot_uniform(shape, rng: np.random.Generator) -> np.ndarray:
    """Keras GlorotUniform: U(-l, l), l = sqrt(6/(fi+fo)); scalar shape has fans (1, 1)."""
    if len(shape) == 0:
        fi = fo = 1
    elif len(shape) == 1:
        fi = fo = shape[0]
    else:
        fi, fo = shape[0], shape[1]
    lim = np.sqrt(6.0 / (fi + fo))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_params(layout: ParamLayout, rng: np.random.Generator, y0_init: Sequence[str] = ()) -> np.ndarray:
    """Random-init flat parameters the way Keras would (kernels Glorot-normal, biases 0)."""
    theta = np.zeros(layout.total, dtype=np.float32)
    for k in range(len(layout.nets)):
        for (w0, w1, fi, fo, b0, b1) in layout.net_slices(k):
            theta[w0:w1] = glorot_normal((fi, fo), rng).reshape(-1)
    for j in range(layout.n_y0):
        kind = y0_init[j] if j < len(y0_init) else "normal"
        theta[layout.y0_offset + j] = (glorot_uniform((), rng) if kind == "uniform" else glorot_normal((), rng))
    return theta


def mlp_forward(theta: torch.Tensor, layout: ParamLayout, k: int, x: torch.Tensor) -> torch.Tensor:
    """Evaluate net k of `layout` on x[..., nin] -> [..., nout]."""
    spec = layout.nets[k]
    sl = layout.net_slices(k)
    h = x
    for li, (w0, w1, fi, fo, b0, b1) in enumerate(sl):
        W = theta[w0:w1].reshape(fi, fo)
        b = theta[b0:b1]
        h = h @ W + b
        if li < len(sl) - 1:
            h = torch.tanh(h) if spec.activation == "tanh" else torch.relu(h)
    return h

"""Keras-equivalent initialisers for the drop-in networks (TF semantics, restated; the reference calls
tf.keras.initializers.GlorotNormal / GlorotUniform at coupledPricing/Networks.py:12-14, coupledMFG/Networks.py:12-16,30-34)."""
from __future__ import annotations

import numpy as np

_TRUNC_STD = 0.87962566103423978
_rng = np.random.default_rng()


def set_seed(seed: int) -> None:
    """Seed the parameter initialiser (the reference never seeds TF; runs there are not reproducible)."""
    global _rng
    _rng = np.random.default_rng(seed)


def _fans(shape):
    if len(shape) == 0:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    return shape[0], shape[1]


def glorot_normal(shape) -> np.ndarray:
    """Truncated normal (|z| <= 2 sigma), stddev sqrt(2 / (fan_in + fan_out)) / 0.8796."""
    fi, fo = _fans(shape)
    std = np.sqrt(2.0 / (fi + fo)) / _TRUNC_STD
    n = int(np.prod(shape)) if len(shape) else 1
    z = _rng.standard_normal(n)
    bad = np.abs(z) > 2.0
    while bad.any():
        z[bad] = _rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2.0
    return (std * z).reshape(shape).astype(np.float32)


def glorot_uniform(shape) -> np.ndarray:
    fi, fo = _fans(shape)
    lim = np.sqrt(6.0 / (fi + fo))
    return _rng.uniform(-lim, lim, size=shape).astype(np.float32)

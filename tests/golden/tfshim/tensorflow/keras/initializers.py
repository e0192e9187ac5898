"""GlorotNormal / GlorotUniform (TF semantics): fans of a scalar shape are (1, 1); the normal is truncated at 2 sigma
and rescaled by 0.87962566103423978."""
import math

import torch

GEN = torch.Generator().manual_seed(1234)


def _fans(shape):
    shape = list(shape)
    if len(shape) == 0:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    return shape[0], shape[1]


class GlorotNormal:
    def __call__(self, shape, dtype=torch.float32):
        fi, fo = _fans(shape)
        std = math.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
        n = 1
        for s in shape:
            n *= int(s)
        z = torch.randn(n, generator=GEN)
        bad = z.abs() > 2
        while bad.any():
            z[bad] = torch.randn(int(bad.sum()), generator=GEN)
            bad = z.abs() > 2
        return (std * z).reshape(list(shape)).to(dtype)


class GlorotUniform:
    def __call__(self, shape, dtype=torch.float32):
        fi, fo = _fans(shape)
        lim = math.sqrt(6.0 / (fi + fo))
        n = 1
        for s in shape:
            n *= int(s)
        return ((torch.rand(n, generator=GEN) * 2 - 1) * lim).reshape(list(shape)).to(dtype)

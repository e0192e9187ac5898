import numpy as np
import torch


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def sqrt(x):
    return torch.sqrt(_t(x))


def abs(x):  # noqa: A001
    return torch.abs(x)


def real(x):
    return torch.real(x)


def lgamma(x):
    return torch.lgamma(x)


def multiply(a, b):
    return a * b


def reduce_std(x, axis=None):
    return x.std(unbiased=False) if axis is None else x.std(dim=axis, unbiased=False)

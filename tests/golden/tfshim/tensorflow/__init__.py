"""Torch-backed stand-in for the small part of the TensorFlow 2.x API that the reference
(ZakariaBensaid/DeepFBSDEJSolvers) uses, so that its source files can be imported and executed UNMODIFIED in an image
without TensorFlow.  TEST INFRASTRUCTURE ONLY (tests/golden/make_golden.py): it exists to pin the oracle against the
reference's own code.  Semantics restated from TF's published definitions: float32 defaults, Dense = act(x @ W + b),
GlorotNormal/GlorotUniform, Keras Adam, tf.abs/tf.maximum sub-gradients, numpy_function = stop-gradient.
Every random draw is logged (`random.LOG`) so the generator script can hand the same noise to the oracle.
"""
import numpy as np
import torch

from . import keras, signal, random, math, nn, dtypes  # noqa: F401

float32 = torch.float32
float64 = torch.float64
newaxis = None

_orig_numpy = torch.Tensor.numpy
torch.Tensor.numpy = lambda self, *a, **k: _orig_numpy(self.detach(), *a, **k)   # tf tensors always have .numpy()


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype if dtype is not None else (torch.float32 if np.asarray(x).dtype.kind == "f" else None))


def ones(shape, dtype=torch.float32):
    return torch.ones(*[int(s) for s in shape], dtype=dtype) if len(shape) else torch.ones((), dtype=dtype)


def zeros(shape, dtype=torch.float32):
    return torch.zeros(*[int(s) for s in shape], dtype=dtype) if len(shape) else torch.zeros((), dtype=dtype)


def ones_like(x):
    return torch.ones_like(_t(x))


def stack(xs, axis=0):
    xs = [_t(x, torch.float32) for x in xs]
    return torch.stack(xs, dim=axis)


def reduce_mean(x, axis=None):
    return x.mean() if axis is None else x.mean(dim=axis)


def reduce_sum(x, axis=None):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis)


def square(x):
    return x * x


def exp(x):
    return torch.exp(_t(x, torch.float32) if not isinstance(x, torch.Tensor) else x)


def sqrt(x):
    return torch.sqrt(_t(x, torch.float32) if not isinstance(x, torch.Tensor) else x)


def abs(x):  # noqa: A001
    return torch.abs(x)


def broadcast_to(x, shape):
    return torch.broadcast_to(x, [int(s) for s in shape])


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def cast(x, dtype):
    return _t(x).to(dtype)


def where(c, a, b):
    a = _t(a, torch.float32) if not isinstance(a, torch.Tensor) else a
    b = _t(b, torch.float32) if not isinstance(b, torch.Tensor) else b
    return torch.where(c, a, b)


def shape(x):
    return list(x.shape)


def maximum(a, b):
    # tf.maximum: the gradient goes to the first argument on ties (x >= y)
    b = _t(b, a.dtype) if not isinstance(b, torch.Tensor) else b
    return torch.where(a >= b, a, b)


def range(n, dtype=torch.float32):  # noqa: A001
    return torch.arange(int(n), dtype=dtype)


def linspace(a, b, n):
    return torch.linspace(float(a), float(b), int(n))


def function(f=None, **kw):
    return f if f is not None else (lambda g: g)


def numpy_function(func, inp, Tout):
    args = [x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x) for x in inp]
    out = func(*args)
    return out.detach() if isinstance(out, torch.Tensor) else torch.as_tensor(np.asarray(out), dtype=Tout)


class _VarTensor(torch.Tensor):
    pass


def Variable(value, trainable=True, dtype=torch.float32, name=None):
    v = _t(value, dtype).clone().detach().requires_grad_(bool(trainable))
    v.assign = lambda new: v.data.copy_(_t(new, dtype))
    return v


class GradientTape:
    LOG = []          # (target value, [grad or None, ...], [variable, ...]) per .gradient() call

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, variables):
        variables = list(variables)
        grads = torch.autograd.grad(target, variables, allow_unused=True)
        GradientTape.LOG.append((float(target.detach()), [None if g is None else g.detach().clone() for g in grads], variables))
        return list(grads)

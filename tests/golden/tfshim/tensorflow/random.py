"""tf.random.* drawn from one seeded torch generator; every draw is appended to LOG as (kind, tensor)."""
import torch

GEN = torch.Generator().manual_seed(0)
LOG = []


def seed(s):
    GEN.manual_seed(int(s))
    LOG.clear()


def normal(shape, mean=0.0, stddev=1.0, dtype=torch.float32):
    x = torch.randn(*[int(s) for s in shape], generator=GEN, dtype=dtype) * stddev + mean
    LOG.append(("normal", x))
    return x


def poisson(shape, lam, dtype=torch.float32):
    lam = lam if isinstance(lam, torch.Tensor) else torch.as_tensor(lam, dtype=torch.float32)
    rate = lam.detach().clamp_min(0).expand(*[int(s) for s in shape], *lam.shape)
    x = torch.poisson(rate.contiguous(), generator=GEN).to(dtype)
    LOG.append(("poisson", x))
    return x


def gamma(shape, alpha, beta=None, dtype=torch.float32):
    conc = torch.full([int(s) for s in shape], float(alpha), dtype=torch.float64)
    x = torch._standard_gamma(conc, generator=GEN)
    if beta is not None:
        x = x / float(beta)
    x = x.to(dtype)
    LOG.append(("gamma", x))
    return x

"""Keras (TF 2.10 OptimizerV2) Adam, restated (TF semantics; not under /root/reference).

Used by the reference through `optimizers.Adam(learning_rate=lRate)`
(coupledPricing/SolversJumpDiff.py:55, coupledMFG/MFGSolvers.py:75):
    t += 1
    alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2)
    theta -= alpha * m / (sqrt(v) + eps)          # eps = 1e-7 on the un-corrected sqrt(v)
This differs from torch.optim.Adam, which applies eps after bias-correcting v.
"""
from __future__ import annotations

import numpy as np
import torch


class KerasAdam:
    def __init__(self, n: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7,
                 dtype=torch.float32):
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(beta1), float(beta2), float(eps)
        self.m = torch.zeros(n, dtype=dtype)
        self.v = torch.zeros(n, dtype=dtype)
        self.t = 0

    def step(self, theta: torch.Tensor, grad: torch.Tensor, mask: torch.Tensor | None = None) -> None:
        """In-place update of `theta`. `mask` (0/1 per parameter) restricts the update to the
        variables handed to `apply_gradients` (the others keep their slots untouched)."""
        self.t += 1
        f32 = np.float32
        if theta.dtype == torch.float32:
            b1p = f32(self.beta1) ** f32(self.t)
            b2p = f32(self.beta2) ** f32(self.t)
            alpha = f32(self.lr) * np.sqrt(f32(1) - b2p) / (f32(1) - b1p)
        else:
            alpha = self.lr * np.sqrt(1 - self.beta2 ** self.t) / (1 - self.beta1 ** self.t)
        with torch.no_grad():
            m_new = self.m + (grad - self.m) * (1 - self.beta1)
            v_new = self.v + (grad * grad - self.v) * (1 - self.beta2)
            upd = float(alpha) * m_new / (torch.sqrt(v_new) + self.eps)
            if mask is not None:
                keep = mask > 0
                self.m = torch.where(keep, m_new, self.m)
                self.v = torch.where(keep, v_new, self.v)
                theta -= torch.where(keep, upd, torch.zeros_like(upd))
            else:
                self.m, self.v = m_new, v_new
                theta -= upd

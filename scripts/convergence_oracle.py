#!/usr/bin/env python
"""The reference-equivalent CPU restatement (oracle/) trained the way scripts/convergence.py trains the CUDA path, for the
`merton_reg` case (SolverGlobalSumLocalReg at the mainMerton.py defaults: 10^4 paths per step, N = 50, lr 3e-4): same initialiser
seed, noise drawn on the CPU the way the reference draws it.  Writes profiles/r2_convergence_oracle_merton_reg.csv (epoch, Y0) -
the curve the CUDA path's Y0-vs-epoch curve is compared with (tests are statistical: the two use different random streams)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle import MertonOracle, KerasAdam, pricing_loss, mlp_forward  # noqa: E402
from oracle.pricing import sample_pricing_noise  # noqa: E402
from oracle.nets import init_params  # noqa: E402


def main():
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    om = MertonOracle(aLin=0.1, limit=30, d=1, **H.MERTON)
    layout = H.pricing_layout("merton", "SumLocalReg", 1)
    theta = torch.tensor(init_params(layout, np.random.default_rng(2026)), requires_grad=True)   # Glorot normal, zero biases
    opt = KerasAdam(layout.total, 3e-4)
    gen = torch.Generator().manual_seed(2026)
    B, t0 = 10000, time.time()
    path = os.path.join(ROOT, "profiles", "r2_convergence_oracle_merton_reg.csv")
    with open(path, "w") as f:
        f.write("epoch,Y0,seconds\n")
        for ep in range(epochs):
            for _ in range(100):
                theta.grad = None
                loss = pricing_loss(om, "SumLocalReg", layout, theta, sample_pricing_noise(om, "SumLocalReg", B, 1, gen), B)
                loss.backward()
                opt.step(theta.data, theta.grad)
            y0 = float(mlp_forward(theta.detach(), layout, 0, torch.tensor([[0.0, 1.0]]))[0, 0])
            f.write("%d,%.8f,%.1f\n" % (ep, y0, time.time() - t0))
            f.flush()


if __name__ == "__main__":
    main()

"""Driver with the command line of coupledMFG/mainMFGComparison.py (reference lines 13-33): the five smart-grid MFG solver
classes on the coupled FBSDE with Cox-process jumps.

    python -m deepfbsdejsolvers_b200.coupledMFG.mainMFGComparison [--nEpochExt 100 --nEpoch 200 --batchSize 128 ...]

The reference script ends by loading `hY0List.csv` / `Y0List.csv` that nothing writes (mainMFGComparison.py:146-147); here
the two files are written (one column per method) and nothing is plotted.  Load curve and constants:
mainMFGComparison.py:83-94, 108.  `--methods`, `--seed` are extra flags.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import (ModelCoupledFBSDE, Net_hat, Net, kerasModels, SolverGlobalFBSDE, SolverMultiStepFBSDE, SolverSumLocalFBSDE,
               SolverGlobalSumLocalReg, SolverGlobalMultiStepReg)

METHODS = ['Global', 'SumMultiStep', 'SumLocal', 'SumLocalReg', 'SumMultiStepReg']

QAverOneDay = np.array([0.26759617, 0.24771933, 0.23588383, 0.221369, 0.21174, 0.2047625, 0.20651067, 0.20098083, 0.20826067, 0.22095067,
                        0.24346833, 0.27283267, 0.3382265, 0.42920433, 0.4875495, 0.50948433, 0.487712, 0.4537295, 0.40911717, 0.3728925,
                        0.347346, 0.3419715, 0.32684, 0.320009, 0.32065767, 0.32586567, 0.31492483, 0.31607417, 0.30411783, 0.29950567,
                        0.307519, 0.33259367, 0.375465, 0.45608333, 0.599178, 0.70970583, 0.7364855, 0.736731, 0.70612667, 0.67284583,
                        0.66692767, 0.64925583, 0.604485, 0.55684567, 0.515597, 0.45097333, 0.3822625, 0.31841833])


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--nbNeuron_hat', type=int, default=20)
    parser.add_argument('--nbNeuron', type=int, default=22)
    parser.add_argument('--nbLayer_hat', type=int, default=2)
    parser.add_argument('--nbLayer', type=int, default=2)
    parser.add_argument('--nEpochExt', type=int, default=100)
    parser.add_argument('--nEpoch', type=int, default=200)
    parser.add_argument('--batchSize', type=int, default=128)
    parser.add_argument('--rafCoef', type=int, default=1)
    parser.add_argument('--jumpFac', type=float, default=2.16)
    parser.add_argument('--nbDays', type=int, default=2)
    parser.add_argument('--lRateY0', type=float, default=0.001)
    parser.add_argument('--lRateLoc', type=float, default=0.00015)
    parser.add_argument('--lRateReg', type=float, default=0.0001)
    parser.add_argument('--couplage', type=str, default='ON')
    parser.add_argument('--jumpModel', type=str, default='stochastic')
    parser.add_argument('--activation_hat', type=str, default="tanh")
    parser.add_argument('--activation', type=str, default="tanh")
    parser.add_argument('--nbSimulation', type=int, default=10**5)
    parser.add_argument('--methods', type=str, default=",".join(METHODS))
    parser.add_argument('--seed', type=int, default=0)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("Args ", args)
    for act in (args.activation_hat, args.activation):
        if act not in ['tanh', 'relu']:
            print(act, 'is invalid. Please choose tanh or relu.')
            sys.exit(0)
    from .. import set_seed
    set_seed(args.seed)
    layerSize_hat = args.nbNeuron_hat * np.ones((args.nbLayer_hat,), dtype=np.int32)
    layerSize = args.nbNeuron * np.ones((args.nbLayer_hat,), dtype=np.int32)       # (sic) mainMFGComparison.py:80
    QAver = np.concatenate([QAverOneDay] * args.nbDays, axis=-1)
    QAver = np.tile(np.expand_dims(QAver, axis=-1), [1, args.rafCoef]).flatten()
    T = float(args.nbDays)
    sig, sig0, theta, h1, h2, A, C, K, R0, S0, alphaTarget, coeffOU, alpha = 0.3, 0.1, 0.12, 0, 600, 150, 80, 50, 2 * 0.12, 0, -0.2, 5., 30
    beta = np.exp(-0.5 * alpha)
    pi, p0, p1, f0, f1 = 0.1, 6.159423723, 87.4286117, 0, 10**4
    mathModel = ModelCoupledFBSDE(T, QAver, R0, args.jumpFac, alpha, beta, coeffOU, A, K, pi, p0, p1, f0, f1, theta, C, S0, h1, h2,
                                  sig0, sig, alphaTarget, args.jumpModel, 1)
    listhY0List, listY0List, names = [], [], []
    for method in [m for m in args.methods.split(",") if m]:
        if method in ['SumMultiStepReg', 'SumLocalReg']:
            kerasModel = kerasModels(Net_hat, Net, method, 1, 1, layerSize_hat, layerSize, args.activation_hat, args.activation)
        elif method in ['SumMultiStep', 'SumLocal']:
            kerasModel = kerasModels(Net_hat, Net, method, 3, 4, layerSize_hat, layerSize, args.activation_hat, args.activation)
        else:
            kerasModel = kerasModels(Net_hat, Net, method, 2, 3, layerSize_hat, layerSize, args.activation_hat, args.activation)
        if method == "Global":
            solver = SolverGlobalFBSDE(mathModel, kerasModel, args.lRateY0, args.couplage, seed=args.seed)
        elif method == "SumMultiStep":
            solver = SolverMultiStepFBSDE(mathModel, kerasModel, args.lRateReg, args.couplage, seed=args.seed)
        elif method == "SumLocal":
            solver = SolverSumLocalFBSDE(mathModel, kerasModel, args.lRateLoc, args.couplage, seed=args.seed)
        elif method == 'SumMultiStepReg':
            solver = SolverGlobalMultiStepReg(mathModel, kerasModel, args.lRateReg, args.couplage, seed=args.seed)
        elif method == 'SumLocalReg':
            solver = SolverGlobalSumLocalReg(mathModel, kerasModel, args.lRateLoc, args.couplage, seed=args.seed)
        else:
            raise ValueError(f"unknown method {method}")
        hY0List, Y0List = solver.train(args.batchSize, args.batchSize * 10, args.nEpoch, args.nEpochExt)
        listhY0List.append(hY0List); listY0List.append(Y0List); names.append(method)
        print('method', method, 'Y0_hat', hY0List[-1], 'Y0', Y0List[-1])
    np.savetxt('hY0List.csv', np.array(listhY0List, dtype=np.float64), delimiter=',', header=",".join(names))
    np.savetxt('Y0List.csv', np.array(listY0List, dtype=np.float64), delimiter=',', header=",".join(names))
    print("wrote hY0List.csv, Y0List.csv (one row per method)")
    return listhY0List, listY0List


if __name__ == "__main__":
    main()

"""Driver with the command line of coupledPricing/mainMerton.py (reference lines 12-25): trains the seven jump-diffusion
solver classes on the Merton European call and compares the learned Y0 with the closed-form series price.

    python -m deepfbsdejsolvers_b200.coupledPricing.mainMerton [--nEpochExt 120 --nEpoch 100 --batchSize 10 ...]

Differences from the reference script: the Y0 / loss curves go to a CSV file instead of a matplotlib window, the
(listY0, duration) tuple returned by `train` is unpacked (mainMerton.py:120-124 plots the tuple), and `--methods`,
`--seed`, `--out` are extra flags.  Model constants: mainMerton.py:57.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import (MertonJumpModel, AbsCoupling, Net, SolverGlobalFBSDE, SolverMultiStepFBSDE1, SolverMultiStepFBSDE2,
               SolverSumLocalFBSDE1, SolverSumLocalFBSDE2, SolverGlobalMultiStepReg, SolverGlobalSumLocalReg)

METHODS = ['Global', 'SumMultiStep1', 'SumMultiStep2', 'SumLocal1', 'SumLocal2', 'SumLocalReg', 'SumMultiStepReg']


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--nbNeuron', type=int, default=21)
    parser.add_argument('--nbLayer', type=int, default=2)
    parser.add_argument('--nEpochExt', type=int, default=120)
    parser.add_argument('--nEpoch', type=int, default=100)
    parser.add_argument('--batchSize', type=int, default=10)
    parser.add_argument('--lRateY0', type=float, default=0.0004)
    parser.add_argument('--lRateLoc', type=float, default=0.0003)
    parser.add_argument('--lRateReg', type=float, default=0.0003)
    parser.add_argument('--activation', type=str, default="tanh")
    parser.add_argument('--aLin', type=float, default=0.1)
    parser.add_argument('--limit', type=int, default=30)
    parser.add_argument('--methods', type=str, default=",".join(METHODS), help="comma-separated subset of the seven methods")
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--out', type=str, default="merton_Y0.csv")
    return parser


def make_solver(method, mathModel, layerSize, activation, lRateY0, lRateLoc, lRateReg, seed):
    bY0, ndimOut = 0, 2
    if method == 'Global':
        bY0, ndimOut = 1, 1
    elif method in ['SumLocalReg', 'SumMultiStepReg']:
        ndimOut = 1
    kerasModelUZ = Net(bY0, ndimOut, layerSize, activation)
    kerasModelGam = Net(0, 1, layerSize, activation)
    if method == "Global":
        return SolverGlobalFBSDE(mathModel, kerasModelUZ, kerasModelGam, lRateY0, seed=seed)
    if method == "SumMultiStep1":
        return SolverMultiStepFBSDE1(mathModel, kerasModelUZ, lRateLoc, seed=seed)
    if method == "SumMultiStep2":
        return SolverMultiStepFBSDE2(mathModel, kerasModelUZ, kerasModelGam, lRateLoc, seed=seed)
    if method == "SumLocal1":
        return SolverSumLocalFBSDE1(mathModel, kerasModelUZ, lRateLoc, seed=seed)
    if method == "SumLocal2":
        return SolverSumLocalFBSDE2(mathModel, kerasModelUZ, kerasModelGam, lRateLoc, seed=seed)
    if method == 'SumMultiStepReg':
        return SolverGlobalMultiStepReg(mathModel, kerasModelUZ, kerasModelGam, lRateReg, seed=seed)
    if method == 'SumLocalReg':
        return SolverGlobalSumLocalReg(mathModel, kerasModelUZ, kerasModelGam, lRateReg, seed=seed)
    raise ValueError(f"unknown method {method}")


def write_csv(path, columns):
    """columns: {name: list}; one row per outer epoch."""
    names = list(columns)
    n = max(len(v) for v in columns.values())
    with open(path, "w") as f:
        f.write("epoch," + ",".join(names) + "\n")
        for i in range(n):
            f.write(str(i) + "," + ",".join(repr(float(columns[k][i])) if i < len(columns[k]) else "" for k in names) + "\n")


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("Args ", args)
    if args.activation not in ['tanh', 'relu']:
        print(args.activation, 'is invalid. Please choose tanh or relu.')
        sys.exit(0)
    from .. import set_seed
    set_seed(args.seed)
    layerSize = args.nbNeuron * np.ones((args.nbLayer,), dtype=np.int32)
    T, N, r, sig, lam, muJ, sigJ, K, x0 = 1, 50, 0.1, 0.3, 3, 0., 0.2, 0.9, 1          # mainMerton.py:57
    func = AbsCoupling(args.aLin)                                                        # mainMerton.py:60-61
    mathModel0 = MertonJumpModel(T, N, r, muJ, sigJ, sig, lam, K, x0, func, args.limit)
    Realprice = mathModel0.A(0, mathModel0.init(1)).numpy()[0]
    print('Merton real price:', Realprice)
    cols = {}
    for method in [m for m in args.methods.split(",") if m]:
        mathModel = MertonJumpModel(T, N, r, muJ, sigJ, sig, lam, K, x0, func, args.limit)
        solver = make_solver(method, mathModel, layerSize, args.activation, args.lRateY0, args.lRateLoc, args.lRateReg, args.seed)
        Y0List, duration = solver.train(args.batchSize, args.batchSize * 10, args.nEpoch, args.nEpochExt)
        print('Y0', Y0List[-1], 'method', method, 'training time %.3f s' % duration)
        cols["Y0_" + method], cols["loss_" + method] = Y0List, solver.lossList
    cols["Y0_closed_formula"] = [Realprice] * args.nEpochExt
    write_csv(args.out, cols)
    print("wrote", args.out)
    return cols


if __name__ == "__main__":
    main()

"""Driver with the command line of coupledPricing/mainVG.py (reference lines 12-24): the seven pure-jump solver classes on
the Variance-Gamma European option against the Lewis/FFT price.

    python -m deepfbsdejsolvers_b200.coupledPricing.mainVG [--nEpochExt 120 --nEpoch 100 --batchSize 10 ...]

CSV output instead of plots; `--methods`, `--seed`, `--out` are extra flags.  Model constants: mainVG.py:54; the trainable
Y0 of the Global solver lives on the Gam network (mainVG.py:91-95).
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import VGmodel, AbsCoupling, Net
from . import SolversPureJump as pj
from .mainMerton import write_csv, METHODS


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--nbNeuron', type=int, default=21)
    parser.add_argument('--nbLayer', type=int, default=2)
    parser.add_argument('--nEpochExt', type=int, default=120)
    parser.add_argument('--nEpoch', type=int, default=100)
    parser.add_argument('--batchSize', type=int, default=10)
    parser.add_argument('--lRateY0', type=float, default=0.0005)
    parser.add_argument('--lRateLoc', type=float, default=0.0003)
    parser.add_argument('--lRateReg', type=float, default=0.00015)
    parser.add_argument('--activation', type=str, default="tanh")
    parser.add_argument('--aLin', type=float, default=0.1)
    parser.add_argument('--methods', type=str, default=",".join(METHODS))
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--out', type=str, default="vg_Y0.csv")
    return parser


def make_solver(method, mathModel, layerSize, activation, lRateY0, lRateLoc, lRateReg, seed):
    kerasModelU = Net(0, 1, layerSize, activation)
    kerasModelGam = Net(1 if method == 'Global' else 0, 1, layerSize, activation)
    if method == "Global":
        return pj.SolverGlobalFBSDE(mathModel, kerasModelU, kerasModelGam, lRateY0, seed=seed)
    if method == "SumMultiStep1":
        return pj.SolverMultiStepFBSDE1(mathModel, kerasModelU, lRateLoc, seed=seed)
    if method == "SumMultiStep2":
        return pj.SolverMultiStepFBSDE2(mathModel, kerasModelU, kerasModelGam, lRateLoc, seed=seed)
    if method == "SumLocal1":
        return pj.SolverSumLocalFBSDE1(mathModel, kerasModelU, lRateLoc, seed=seed)
    if method == "SumLocal2":
        return pj.SolverSumLocalFBSDE2(mathModel, kerasModelU, kerasModelGam, lRateLoc, seed=seed)
    if method == 'SumMultiStepReg':
        return pj.SolverGlobalMultiStepReg(mathModel, kerasModelU, kerasModelGam, lRateReg, seed=seed)
    if method == 'SumLocalReg':
        return pj.SolverGlobalSumLocalReg(mathModel, kerasModelU, kerasModelGam, lRateReg, seed=seed)
    raise ValueError(f"unknown method {method}")


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("Args ", args)
    if args.activation not in ['tanh', 'relu']:
        print(args.activation, 'is invalid. Please choose tanh or relu.')
        sys.exit(0)
    from .. import set_seed
    set_seed(args.seed)
    layerSize = args.nbNeuron * np.ones((args.nbLayer,), dtype=np.int32)
    T, N, r, theta, kappa, sigmaJ, K, x0 = 1, 30, 0.1, -0.1, 0.1, 0.2, 1, 1               # mainVG.py:54
    func = AbsCoupling(args.aLin)
    mathModel0 = VGmodel(T, N, r, theta, kappa, sigmaJ, K, x0, func)
    Realprice = mathModel0.A(0, mathModel0.init(1)).numpy()[0]
    print('VG real price:', Realprice)
    cols = {}
    for method in [m for m in args.methods.split(",") if m]:
        mathModel = VGmodel(T, N, r, theta, kappa, sigmaJ, K, x0, func)
        solver = make_solver(method, mathModel, layerSize, args.activation, args.lRateY0, args.lRateLoc, args.lRateReg, args.seed)
        Y0List, durations = solver.train(args.batchSize, args.batchSize * 10, args.nEpoch, args.nEpochExt)
        print('Y0', Y0List[-1], 'method', method)
        cols["Y0_" + method], cols["loss_" + method] = Y0List, solver.lossList
    cols["Y0_closed_formula"] = [Realprice] * args.nEpochExt
    write_csv(args.out, cols)
    print("wrote", args.out)
    return cols


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""ms per training iteration of the two-network jump schemes with the jump network on FFMA tiles vs tcgen05 (jump_tc.cuh),
mainMerton.py / mainVG.py shapes (M = 5000) at several batch sizes."""
import json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="", help="run one case only (for ncu), e.g. 'vg Global'")
    ap.add_argument("--paths", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--ffma", action="store_true", help="with --case: time the fp32 FFMA tile kernels instead")
    a = ap.parse_args()
    import helpers as H
    from bench_configs import time_solver
    from deepfbsdejsolvers_b200 import Context, set_seed
    from deepfbsdejsolvers_b200 import coupledPricing as cp
    from deepfbsdejsolvers_b200.coupledPricing import SolversPureJump as pj
    ctx = Context.default(0)
    set_seed(0)
    M, V = H.MERTON, H.VG
    net = lambda bY0, nout: cp.Net(bY0, nout, [21, 21], "tanh")
    merton = lambda: cp.MertonJumpModel(M["T"], M["N"], M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(0.1), 30)
    vg = lambda: cp.VGmodel(V["T"], V["N"], V["r"], V["theta"], V["kappa"], V["sigmaJ"], V["K"], V["x0"], cp.AbsCoupling(0.1))
    merton10 = lambda: cp.MertonJumpModel(M["T"], 100, M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(0.1), 100, d=10)
    cases = [("merton10 Global", lambda tc: cp.SolverGlobalFBSDE(merton10(), net(1, 10), net(0, 1), 4e-4, M=256, tensor_cores=tc)),
             ("merton Global", lambda tc: cp.SolverGlobalFBSDE(merton(), net(1, 1), net(0, 1), 4e-4, tensor_cores=tc)),
             ("merton SumLocal2", lambda tc: cp.SolverSumLocalFBSDE2(merton(), net(0, 2), net(0, 1), 3e-4, tensor_cores=tc)),
             ("merton MultiStep1", lambda tc: cp.SolverMultiStepFBSDE1(merton(), net(0, 2), 3e-4, tensor_cores=tc)),
             ("vg SumLocal1", lambda tc: pj.SolverSumLocalFBSDE1(vg(), net(0, 1), 3e-4, tensor_cores=tc)),
             ("vg Global", lambda tc: pj.SolverGlobalFBSDE(vg(), net(0, 1), net(1, 1), 5e-4, tensor_cores=tc)),
             ("vg MultiStep2", lambda tc: pj.SolverMultiStepFBSDE2(vg(), net(0, 1), net(0, 1), 3e-4, tensor_cores=tc))]
    if a.case:
        print(json.dumps({"case": a.case, "paths": a.paths,
                          "ffma_ms" if a.ffma else "tcgen05_ms": time_solver(dict(cases)[a.case](not a.ffma), a.paths, a.iters or 3, ctx)}))
        return
    for name, mk in cases:
        if name.startswith("merton10"):
            continue                              # d = 10: run with --case (B = 2^16 takes a while on the FFMA tiles)
        for B in (10, 100, 1000):
            iters = 100 if B <= 100 else 20
            r = {"case": name, "paths": B}
            for tc in (False, True):
                r["tcgen05_ms" if tc else "ffma_ms"] = time_solver(mk(tc), B, iters, ctx)
            r["speedup"] = r["ffma_ms"] / r["tcgen05_ms"]
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()

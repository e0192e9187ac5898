"""GPU parity: the fused sm_100a kernels (through the C-ABI) against the CPU oracle on identical injected noise.

Bar (BASELINE.json north_star): losses, Y0 / Y and Z trajectories within 1e-5 relative in fp32.  Gradients are
compared against the float64 oracle and must be as close to it as fp32 arithmetic allows (the fp32 oracle's own
distance to float64 is the yardstick)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import MertonOracle, VGOracle, MFGOracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(a, b, rtol=RTOL, atol=2e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert np.all(err <= tol), f"max err {err.max():.3e} (rel {np.max(err / (np.abs(b) + 1e-30)):.3e})"


def _grad_check(g_gpu, g64, g32):
    scale = np.abs(g64).max() + 1e-30
    e_gpu = np.abs(g_gpu - g64).max() / scale
    e_32 = np.abs(g32 - g64).max() / scale
    # entries such as the output bias of the jump network are exact zeros in exact arithmetic (gam - mean(comp)):
    # what is left there is fp32 summation noise of either implementation, hence a bound relative to max|g|
    assert e_gpu <= max(1e-4, 10.0 * e_32), f"gradient error {e_gpu:.3e} vs fp32-oracle error {e_32:.3e}"


@pytest.mark.parametrize("scheme", ["Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2", "SumLocalReg", "MultiStepReg"])
@pytest.mark.parametrize("B", [10, 37])
def test_merton_d1(ctx, scheme, B):
    M = 160
    om = MertonOracle(aLin=H.ALIN, limit=30, d=1, **H.MERTON)
    layout = H.pricing_layout("merton", scheme, 1)
    theta = H.random_theta(layout, 1)
    noise = H.merton_noise(om, B, M, seed=5, with_jmc=not scheme.endswith("Reg"))
    l32, g32, aux = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", H.MERTON, scheme, layout, d=1, M=0 if scheme.endswith("Reg") else M)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]) if "JMC" in noise else None)
    out, tx, ty, tz = s.loss(B, traj=True)
    _close(out[0], l64)
    _close(tx[:, 0, :], aux64["X"][:, :, 0])
    _close(ty[:aux64["Y"].shape[0]], aux64["Y"])
    if "Z" in aux64:
        _close(tz[:, 0, :], aux64["Z"][:, :, 0], atol=5e-6)
    g = s.grad(B)
    _close(g[0], l64)
    _grad_check(g[4:], g64, g32)


@pytest.mark.parametrize("scheme", ["Global", "SumLocal1", "SumLocalReg"])
@pytest.mark.parametrize("tensor_cores", [False, True])
def test_merton_d1_price_table(ctx, scheme, tensor_cores):
    """d = 1 with the closed form A(i, X) read from the per-step Hermite table (the model class's default) instead of the
    30-term series: same parity bar."""
    B, M = 300, 160
    om = MertonOracle(aLin=H.ALIN, limit=30, d=1, **H.MERTON)
    layout = H.pricing_layout("merton", scheme, 1)
    theta = H.random_theta(layout, 11)
    reg = scheme.endswith("Reg")
    noise = H.merton_noise(om, B, M, seed=12, with_jmc=not reg)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", H.MERTON, scheme, layout, d=1, M=0 if reg else M, price_table=True, tensor_cores=tensor_cores)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]) if "JMC" in noise else None)
    out, tx, ty, tz = s.loss(B, traj=True)
    _close(out[0], l64, rtol=2e-5 if tensor_cores else RTOL)
    _close(tx[:, 0, :], aux64["X"][:, :, 0])
    g = s.grad(B)
    scale = np.abs(g64).max() + 1e-30
    e_gpu = np.abs(g[4:] - g64).max() / scale
    assert e_gpu <= (3e-4 if tensor_cores else 1e-4), f"gradient error {e_gpu:.3e}"


@pytest.mark.parametrize("scheme", ["Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2", "SumLocalReg", "MultiStepReg"])
@pytest.mark.parametrize("price_table", [False, True])
def test_merton_d10(ctx, scheme, price_table):
    B, M, d = 64, 96, 10
    p = dict(H.MERTON, N=20)
    om = MertonOracle(aLin=H.ALIN, limit=100, d=d, **p)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 2)
    noise = H.merton_noise(om, B, M, seed=6, with_jmc=not scheme.endswith("Reg"))
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=0 if scheme.endswith("Reg") else M, limit=100,
                         price_table=price_table)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]) if "JMC" in noise else None)
    out, tx, ty, tz = s.loss(B, traj=True)
    _close(out[0], l64)
    _close(tx, aux64["X"].transpose(0, 2, 1))
    _close(ty[:aux64["Y"].shape[0]], aux64["Y"])
    g = s.grad(B)
    _grad_check(g[4:], g64, g32)


@pytest.mark.parametrize("scheme", ["Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2", "SumLocalReg", "MultiStepReg"])
def test_vg(ctx, scheme):
    B, M = 24, 128
    om = VGOracle(aLin=H.ALIN, **H.VG)
    layout = H.pricing_layout("vg", scheme, 1)
    theta = H.random_theta(layout, 3)
    noise = H.vg_noise(om, B, M, seed=7, with_jmc=not scheme.endswith("Reg"))
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "vg", H.VG, scheme, layout, M=0 if scheme.endswith("Reg") else M)
    s.set_theta(theta)
    s.set_noise(B, None, H.to_planes(noise["J"]), H.to_planes(noise["JMC"]) if "JMC" in noise else None)
    out, tx, ty, _ = s.loss(B, traj=True)
    _close(out[0], l64)
    _close(tx[:, 0, :], aux64["X"][:, :, 0])
    _close(ty[:aux64["Y"].shape[0]], aux64["Y"])
    g = s.grad(B)
    _grad_check(g[4:], g64, g32)


@pytest.mark.parametrize("scheme", ["Global", "MultiStep", "SumLocal", "SumLocalReg", "MultiStepReg"])
@pytest.mark.parametrize("jumpModel", ["stochastic", "constant"])
def test_mfg(ctx, scheme, jumpModel):
    from oracle.mfg import sample_mfg_noise
    B = 130
    p = H.mfg_params(1, jumpModel)
    om = MFGOracle(**p)
    layout = H.mfg_layout(scheme)
    theta = H.random_theta(layout, 4)
    noise = sample_mfg_noise(om, B, torch.Generator().manual_seed(8))
    (lh32, li32), g32, _ = H.oracle_mfg(om, scheme, layout, theta, noise, B)
    (lh64, li64), g64, aux64 = H.oracle_mfg(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_mfg(ctx, p, scheme, layout)
    s.set_theta(theta)
    s.set_noise(B, noise["dW0"].numpy(), noise["dW"].numpy(), noise["dN"].numpy())
    out, tx, ty, _ = s.loss(B, traj=True)
    _close(out[1], lh64, rtol=2e-5)
    _close(out[2], li64, rtol=2e-5)
    _close(out[0], lh64 + li64, rtol=2e-5)
    _close(tx[:, 0, :], aux64["hS"], atol=1e-5)
    _close(tx[:, 1, :], aux64["S"], atol=1e-5)
    n = aux64["hY"].shape[0]
    _close(ty[:n, 0, :], aux64["hY"], rtol=2e-5, atol=2e-4)
    _close(ty[:n, 1, :], aux64["Y"], rtol=2e-5, atol=2e-4)
    g = s.grad(B)
    _grad_check(g[4:], g64, g32)


def test_price_kernel_matches_closed_form(ctx):
    """MertonJumpModel.A / VGmodel.A through the C-ABI vs the reference's known answers."""
    from deepfbsdejsolvers_b200.coupledPricing import MertonJumpModel, VGmodel, AbsCoupling
    m = H.MERTON
    mm = MertonJumpModel(m["T"], m["N"], m["r"], m["muJ"], m["sigmaJ"], m["sigma"], m["lam"], m["K"], m["x0"], AbsCoupling(0.1), 30)
    assert abs(float(mm.A(0, mm.init(1)).numpy()[0]) - 0.2714569268) < 2e-6
    _close(mm.A(25, np.array([0.8, 1.0, 1.2], dtype=np.float32)).numpy(), [0.07911842, 0.20222704, 0.36822489], rtol=2e-5)
    _close(mm.A(49, np.array([0.8, 1.0, 1.2], dtype=np.float32)).numpy(), [0.00217845, 0.10377461, 0.30217749], rtol=2e-5)
    for tab in (False, True):
        m10 = MertonJumpModel(m["T"], 100, m["r"], m["muJ"], m["sigmaJ"], m["sigma"], m["lam"], m["K"], m["x0"], AbsCoupling(0.1), 100,
                              d=10, price_table=tab)
        assert abs(float(m10.A(0, m10.init(1)).numpy()[0]) - 0.1109224) < 2e-6
        om = MertonOracle(aLin=0.1, limit=100, d=10, T=1.0, N=100, r=0.1, muJ=0.0, sigmaJ=0.2, sigma=0.3, lam=3.0, K=0.9, x0=1.0)
        X = torch.exp(0.25 * torch.randn(512, 10, generator=torch.Generator().manual_seed(1)))
        for i in (0, 37, 98, 99):
            om.dtype = torch.float64
            ref = om.A(i, X.double()).numpy()
            _close(m10.A(i, X).numpy(), ref, rtol=1e-5, atol=3e-6)
    v = H.VG
    vm = VGmodel(v["T"], v["N"], v["r"], v["theta"], v["kappa"], v["sigmaJ"], v["K"], v["x0"], AbsCoupling(0.1))
    assert abs(float(vm.A(0, vm.init(1)).numpy()[0]) - 0.1331402194) < 2e-6
    _close(vm.A(15, np.array([0.9, 1.0, 1.1], dtype=np.float32)).numpy(), [0.02925149, 0.08252713, 0.16135454], rtol=2e-5)
    assert abs(vm.correction - (-0.0796816965)) < 1e-9


def test_adam_matches_keras_form(ctx):
    from oracle import KerasAdam
    om = MertonOracle(aLin=H.ALIN, limit=30, d=1, **H.MERTON)
    layout = H.pricing_layout("merton", "SumLocalReg", 1)
    theta = H.random_theta(layout, 9)
    s = H.native_pricing(ctx, "merton", H.MERTON, "SumLocalReg", layout)
    s.set_theta(theta)
    rng = np.random.default_rng(0)
    th = torch.tensor(theta.copy())
    opt = KerasAdam(layout.total, 3e-4)
    for _ in range(5):
        g = rng.standard_normal(layout.total).astype(np.float32) * np.float32(1e-2)
        opt.step(th, torch.tensor(g))
        s.adam_step(3e-4, grad=ctx.to_device(g))
    _close(s.get_theta(), th.numpy(), rtol=1e-6, atol=1e-8)


def test_sparse_jump_injection_equals_dense(ctx):
    """fbsdej_solver_set_noise_sparse_jumps (jump planes as their non-zero entries) reproduces set_noise bit for bit."""
    d, B, scheme = 10, 500, "SumLocalReg"
    p = dict(H.MERTON, N=8)
    om = MertonOracle(aLin=H.ALIN, limit=100, d=d, **p)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 3)
    noise = H.merton_noise(om, B, 0, seed=4, with_jmc=False)
    dW, J = H.to_planes(noise["dW"]), H.to_planes(noise["J"])
    assert 0 < np.count_nonzero(J) < 0.5 * J.size
    outs = []
    for sparse in (False, True):
        for tc in (False, True):
            s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, limit=100, tensor_cores=tc)
            s.set_theta(theta)
            if sparse:
                s.set_noise_sparse_jumps(B, dW, J)
            else:
                s.set_noise(B, dW, J, None)
            outs.append(s.grad(B).copy())
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])


# ---- wider networks (mainMerton.py:13-14 / mainMFGComparison.py:14-17 expose nbNeuron; SURVEY 8d names an H = 32 variant of
# config 3): fp32 FFMA kernels compiled for padded widths 32 (H <= 31) and 36 (H <= 35, compensator-free solvers) -----------------
@pytest.mark.parametrize("scheme,d,Hn,B,M", [("SumLocalReg", 10, 32, 200, 0), ("MultiStepReg", 1, 35, 150, 0), ("SumLocalReg", 10, 27, 140, 0),
                                              ("Global", 1, 30, 12, 90), ("SumLocal1", 1, 28, 20, 60), ("MultiStep2", 10, 26, 40, 48)])
def test_wider_hidden_layers(ctx, scheme, d, Hn, B, M):
    p = dict(H.MERTON, N=6)
    om = MertonOracle(aLin=H.ALIN, limit=30 if d == 1 else 100, d=d, **p)
    layout = H.pricing_layout("merton", scheme, d, H=Hn)
    theta = H.random_theta(layout, 71)
    noise = H.merton_noise(om, B, max(M, 1), seed=72, with_jmc=M > 0)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=M, limit=30 if d == 1 else 100)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]) if M > 0 else None)
    out, tx, ty, tz = s.loss(B, traj=True)
    _close(out[0], l64)
    _close(tx, aux64["X"].transpose(0, 2, 1))
    g = s.grad(B)
    _close(g[0], l64)
    _grad_check(g[4:], g64, g32)
    # three training steps through the CUDA graph run and move the parameters
    s.train_steps(1, B, 3, 1e-3)
    ctx.sync()
    assert np.isfinite(s.get_theta()).all() and np.abs(s.get_theta() - theta).max() > 0


def test_wider_hidden_layers_mfg(ctx):
    from oracle.mfg import sample_mfg_noise
    p = H.mfg_params(1, "stochastic")
    om = MFGOracle(**p)
    layout = H.mfg_layout("SumLocal", Hh=28, H=31)
    theta = H.random_theta(layout, 73)
    B = 70
    noise = sample_mfg_noise(om, B, torch.Generator().manual_seed(74))
    (lh32, li32), g32, _ = H.oracle_mfg(om, "SumLocal", layout, theta, noise, B)
    (lh64, li64), g64, aux64 = H.oracle_mfg(om, "SumLocal", layout, theta, noise, B, dtype=torch.float64)
    s = H.native_mfg(ctx, p, "SumLocal", layout)
    s.set_theta(theta)
    s.set_noise(B, noise["dW0"].numpy(), noise["dW"].numpy(), noise["dN"].numpy())
    g = s.grad(B)
    _close(g[1], lh64)
    _close(g[2], li64)
    _grad_check(g[4:], g64, g32)


def test_unsupported_widths_fail_loudly(ctx):
    from deepfbsdejsolvers_b200 import FbsdejError
    p = dict(H.MERTON, N=4)
    with pytest.raises(FbsdejError):       # H = 36 is not compiled
        H.native_pricing(ctx, "merton", p, "SumLocalReg", H.pricing_layout("merton", "SumLocalReg", 1, H=36), d=1)
    with pytest.raises(FbsdejError):       # jump schemes stop at H = 31
        H.native_pricing(ctx, "merton", p, "Global", H.pricing_layout("merton", "Global", 1, H=33), d=1, M=8)


# ---- other depths (the reference's --nbLayer, mainMerton.py:13 / mainVG.py:13 / mainMFGComparison.py:14-15): 1 or 3 equal hidden
# layers on the fp32 FFMA kernels ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [1, 3])
@pytest.mark.parametrize("kind,scheme,d,B,M", [("merton", "Global", 1, 10, 150), ("merton", "Global", 1, 37, 60), ("merton", "SumLocal1", 1, 24, 80),
                                               ("merton", "MultiStep2", 10, 40, 48), ("merton", "SumLocalReg", 10, 300, 0),
                                               ("merton", "MultiStepReg", 1, 200, 0), ("vg", "MultiStep2", 1, 24, 100), ("vg", "SumLocal2", 1, 12, 130)])
def test_other_network_depths(ctx, L, kind, scheme, d, B, M):
    if kind == "merton":
        p = dict(H.MERTON, N=7)
        om = MertonOracle(aLin=H.ALIN, limit=30 if d == 1 else 100, d=d, **p)
        noise = H.merton_noise(om, B, max(M, 1), seed=82, with_jmc=M > 0)
    else:
        p = dict(H.VG, N=7)
        om = VGOracle(aLin=H.ALIN, **p)
        noise = H.vg_noise(om, B, max(M, 1), seed=82, with_jmc=M > 0)
    layout = H.pricing_layout(kind, scheme, d, L=L)
    theta = H.random_theta(layout, 81)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, kind, p, scheme, layout, d=d, M=M, limit=30 if d == 1 else 100)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]) if "dW" in noise else None, H.to_planes(noise["J"]), H.to_planes(noise["JMC"]) if M > 0 else None)
    out, tx, ty, tz = s.loss(B, traj=True)
    _close(out[0], l64)
    _close(tx, aux64["X"].transpose(0, 2, 1))
    _close(ty[:aux64["Y"].shape[0]], aux64["Y"])
    g = s.grad(B)
    _close(g[0], l64)
    _grad_check(g[4:], g64, g32)
    y = s.net_forward(0, np.array([[0.0] + [1.0] * d], dtype=np.float32))          # Y0 report path: Net.call with the same depth
    from oracle import mlp_forward
    y_ref = mlp_forward(torch.tensor(theta, dtype=torch.float64), layout, 0, torch.tensor([[0.0] + [1.0] * d], dtype=torch.float64))
    assert np.abs(y - y_ref.numpy()).max() < 5e-6
    s.train_steps(1, B, 3, 1e-3)
    ctx.sync()
    assert np.isfinite(s.get_theta()).all() and np.abs(s.get_theta() - theta).max() > 0


@pytest.mark.parametrize("L", [1, 3])
@pytest.mark.parametrize("scheme", ["Global", "SumLocal", "MultiStepReg"])
def test_other_network_depths_mfg(ctx, L, scheme):
    from oracle.mfg import sample_mfg_noise
    p = H.mfg_params(1, "stochastic")
    om = MFGOracle(**p)
    layout = H.mfg_layout(scheme, L=L)
    theta = H.random_theta(layout, 83)
    B = 70
    noise = sample_mfg_noise(om, B, torch.Generator().manual_seed(84))
    (lh32, li32), g32, _ = H.oracle_mfg(om, scheme, layout, theta, noise, B)
    (lh64, li64), g64, aux64 = H.oracle_mfg(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_mfg(ctx, p, scheme, layout)
    s.set_theta(theta)
    s.set_noise(B, noise["dW0"].numpy(), noise["dW"].numpy(), noise["dN"].numpy())
    g = s.grad(B)
    _close(g[1], lh64)
    _close(g[2], li64)
    _grad_check(g[4:], g64, g32)


def test_drop_in_classes_take_nbLayer(ctx):
    """Net(bY0, ndimOut, layerSize * np.ones(nbLayer), activation) as mainMerton.py:88,101 builds it, for nbLayer = 1 and 3."""
    from deepfbsdejsolvers_b200 import coupledPricing as cp, set_seed
    set_seed(3)
    M = H.MERTON
    for nbLayer in (1, 3):
        mm = cp.MertonJumpModel(M["T"], 6, M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(H.ALIN), 30)
        layer = 21 * np.ones((nbLayer,), dtype=np.int32)
        solver = cp.SolverSumLocalFBSDE2(mm, cp.Net(0, 2, layer, "tanh"), cp.Net(0, 1, layer, "tanh"), 3e-4, M=64, ctx=ctx)
        listY0, _ = solver.train(8, 16, 4, 2)
        assert len(listY0) == 2 and np.isfinite(listY0).all() and np.isfinite(solver.lossList).all()
        assert solver.native.nets[0].L == nbLayer

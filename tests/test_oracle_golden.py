"""CPU: pins the oracle (oracle/) against
  (a) tests/golden/*.npz - outputs of the reference's OWN source files executed unmodified through tests/golden/tfshim
      (generator: tests/golden/make_golden.py): loss, tape gradients, post-Adam parameters and reported Y0 of one
      training step of each of its 19 solver classes, on the increments the reference drew;
  (b) the closed-form known answers embedded in the reference."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import helpers as H
from oracle import KerasAdam, MertonOracle, VGOracle, MFGOracle, mlp_forward, pricing_loss, mfg_loss

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_case(path):
    z = np.load(path, allow_pickle=False)
    return {k: (z[k].item() if z[k].shape == () else z[k]) for k in z.files}


def oracle_of(case, dtype=torch.float32):
    kind, scheme = str(case["kind"]), str(case["scheme"])
    if kind == "merton":
        om = MertonOracle(T=case["T"], N=int(case["N"]), r=case["r"], muJ=case["muJ"], sigmaJ=case["sigmaJ"], sigma=case["sigma"],
                          lam=case["lam"], K=case["K"], x0=case["x0"], aLin=0.1, limit=30, d=1, dtype=dtype)
        layout = H.pricing_layout("merton", scheme, 1)
    elif kind == "vg":
        om = VGOracle(T=case["T"], N=int(case["N"]), r=case["r"], theta=case["theta"], kappa=case["kappa"], sigmaJ=case["sigmaJ"],
                      K=case["K"], x0=case["x0"], aLin=0.1, dtype=dtype)
        layout = H.pricing_layout("vg", scheme, 1)
    else:
        keys = ("T", "R0", "jumpFactor", "alpha", "beta", "coeffOU", "A", "K", "pi", "p0", "p1", "f0", "f1", "theta", "C", "S0", "h1",
                "h2", "sig0", "sig", "alphaTarget", "coeffEqui")
        om = MFGOracle(QAver=case["QAver"], jumpModel="stochastic", dtype=dtype, **{k: case[k] for k in keys})
        layout = H.mfg_layout(scheme)
    return om, layout


def noise_of(case, dtype=torch.float32):
    t = lambda a: torch.tensor(a, dtype=dtype)
    if str(case["kind"]) == "mfg":
        return {"dW0": t(case["dW0"]), "dW": t(case["dW"]), "dN": t(case["dN"])}
    nz = {"J": t(case["J"])[..., None]}
    if "dW" in case:
        nz["dW"] = t(case["dW"])[..., None]
    if "JMC" in case:
        nz["JMC"] = t(case["JMC"])[..., None]
    return nz


def eval_oracle(case, dtype=torch.float32):
    om, layout = oracle_of(case, dtype)
    th = torch.tensor(case["theta0"], dtype=dtype, requires_grad=True)
    nz = noise_of(case, dtype)
    B = int(case["B"])
    if str(case["kind"]) == "mfg":
        lh, li = mfg_loss(om, str(case["scheme"]), layout, th, nz, B)
        loss = lh + li
    else:
        loss = pricing_loss(om, str(case["scheme"]), layout, th, nz, B)
    loss.backward()
    return om, layout, float(loss.detach()), th.grad.detach()


CASES = sorted(glob.glob(os.path.join(GOLD, "*.npz")))


def test_fixtures_present():
    assert len(CASES) == 19, "run tests/golden/make_golden.py (needs /root/reference)"


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_reproduces_reference_step(path):
    case = load_case(path)
    om, layout, loss, grad = eval_oracle(case)
    assert layout.total == case["theta0"].size
    assert abs(loss - case["loss"]) <= 2e-5 * abs(case["loss"]), (loss, case["loss"])
    g_ref = case["grad"].astype(np.float64)
    err = np.abs(grad.numpy().astype(np.float64) - g_ref).max() / np.abs(g_ref).max()
    assert err < 5e-5, f"gradient: rel-to-max error {err:.2e}"
    # Keras-form Adam on the REFERENCE gradient must land on the reference's post-update parameters
    th = torch.tensor(case["theta0"].copy())
    KerasAdam(layout.total, float(case["lr"])).step(th, torch.tensor(case["grad"]))
    np.testing.assert_allclose(th.numpy(), case["theta1"], rtol=0, atol=2e-7)
    # ... and the oracle's own gradient gives the same step wherever the gradient is not rounding noise
    th2 = torch.tensor(case["theta0"].copy())
    KerasAdam(layout.total, float(case["lr"])).step(th2, grad)
    solid = np.abs(g_ref) > 1e-3 * np.abs(g_ref).max()
    np.testing.assert_allclose(th2.numpy()[solid], case["theta1"][solid], rtol=0, atol=float(case["lr"]) * 2e-2)
    # reported Y0: trainable scalar (Global) or net(0, x0)[0] with the updated parameters
    t1 = torch.tensor(case["theta1"])
    scheme = str(case["scheme"])
    if str(case["kind"]) == "mfg":
        if scheme == "Global":
            y0h, y0 = float(t1[layout.y0_offset]), float(t1[layout.y0_offset + 1])
        else:
            st = om.init(1)
            y0h = float(mlp_forward(t1, layout, 0, om.proj_states(st))[0, 0])
            y0 = float(mlp_forward(t1, layout, 1, om.all_states(st))[0, 0])
        assert abs(y0h - case["Y0_hat_report"]) < 2e-6 and abs(y0 - case["Y0_report"]) < 2e-6
    else:
        if scheme == "Global":
            y0 = float(t1[layout.y0_offset])
        else:
            y0 = float(mlp_forward(t1, layout, 0, torch.tensor([[0.0, float(case["x0"])]]))[0, 0])
        assert abs(y0 - case["Y0_report"]) < 2e-6


OFF_CASES = sorted(glob.glob(os.path.join(GOLD, "off", "*.npz")))


def off_masks(layout, scheme):
    """Variables handed to apply_gradients in the two phases of couplage 'OFF' (MFGSolvers.py:50-64)."""
    n0 = layout.offsets[1]
    hat, ind = np.zeros(layout.total, np.float32), np.zeros(layout.total, np.float32)
    hat[:n0] = 1
    ind[n0:layout.y0_offset] = 1
    if scheme == "Global":
        hat[layout.y0_offset] = 1
        ind[layout.y0_offset + 1] = 1
    return hat, ind


@pytest.mark.parametrize("path", OFF_CASES, ids=[os.path.basename(p)[:-4] for p in OFF_CASES])
def test_oracle_reproduces_reference_couplage_off(path):
    """couplage = 'OFF' (MFGSolvers.py:92-115): a step on the projected player's loss over model_hat's variables, then a step on
    the individual player's loss over model's variables with the SAME optimizer object - the step counter carries over (t = 2
    in the second bias correction) while the slots of the second group start from zero."""
    c = load_case(path)
    assert len(OFF_CASES) == 2
    scheme, B = str(c["scheme"]), int(c["B"])
    om, layout = oracle_of(c)
    hat, ind = off_masks(layout, scheme)
    th = torch.tensor(c["theta0"].copy())
    opt = KerasAdam(layout.total, float(c["lr"]))
    for ph, mask, pick in ((1, hat, 0), (2, ind, 1)):
        t = th.clone().requires_grad_(True)
        nz = {k: torch.tensor(c[f"p{ph}_{k}"]) for k in ("dW0", "dW", "dN")}
        loss = mfg_loss(om, scheme, layout, t, nz, B)[pick]
        loss.backward()
        assert abs(float(loss.detach()) - c[f"p{ph}_loss"]) <= 2e-5 * abs(c[f"p{ph}_loss"])
        g_ref = c[f"p{ph}_grad"].astype(np.float64)
        err = np.abs(t.grad.numpy() * mask - g_ref).max() / np.abs(g_ref).max()
        assert err < 5e-5, f"phase {ph} gradient: rel-to-max error {err:.2e}"
        opt.step(th, torch.tensor(c[f"p{ph}_grad"]), torch.tensor(mask))
    np.testing.assert_allclose(th.numpy(), c["theta2"], rtol=0, atol=2e-7)
    # with a fresh optimizer for the second phase (t = 1 instead of 2) the second group's update differs: the carry-over matters
    th_bad = torch.tensor(c["theta0"].copy())
    KerasAdam(layout.total, float(c["lr"])).step(th_bad, torch.tensor(c["p1_grad"]), torch.tensor(hat))
    KerasAdam(layout.total, float(c["lr"])).step(th_bad, torch.tensor(c["p2_grad"]), torch.tensor(ind))
    assert np.abs(th_bad.numpy() - c["theta2"]).max() > 1e-5


TRAJ = os.path.join(GOLD, "traj", "merton_Global_25steps.npz")


def test_oracle_follows_the_reference_training_trajectory():
    """25 consecutive Adam steps of the reference's own SolverGlobalFBSDE (Merton, its 5000 compensator samples) on the recorded
    increments: the oracle's loss and the trainable Y0 after every step against what the reference computed on the way."""
    c = load_case(TRAJ)
    om, layout = oracle_of(c)
    B, n = int(c["B"]), int(c["nsteps"])
    th = torch.tensor(c["theta0"].copy())
    opt = KerasAdam(layout.total, float(c["lr"]))
    for k in range(n):
        t = th.clone().requires_grad_(True)
        nz = {key: torch.tensor(c[key][k])[..., None] for key in ("dW", "J", "JMC")}
        loss = pricing_loss(om, "Global", layout, t, nz, B)
        loss.backward()
        assert abs(float(loss.detach()) - c["losses"][k]) <= 3e-5 * abs(c["losses"][k]), (k, float(loss.detach()), c["losses"][k])
        opt.step(th, t.grad)
        assert abs(float(th[layout.y0_offset]) - c["Y0_after_step"][k]) <= 3e-6, (k, float(th[layout.y0_offset]), c["Y0_after_step"][k])
    solid = np.abs(c["theta_final"] - c["theta0"]) > 0.2 * n * float(c["lr"])       # entries whose gradient kept its sign
    np.testing.assert_allclose(th.numpy()[solid], c["theta_final"][solid], rtol=0, atol=0.03 * n * float(c["lr"]))


TRAJ_REG = {k: os.path.join(GOLD, "traj", f"{k}_SumLocalReg_300steps.npz") for k in ("merton", "vg")}


def reg_trajectory_inputs(c):
    """The seeded increments of a long Reg trajectory (golden/noise_streams.py), checked against the fixture's checksums:
    (dW or None, J), float32 [nsteps, N, B]."""
    sys.path.insert(0, GOLD)
    import noise_streams
    n, N, B, dt = int(c["nsteps"]), int(c["N"]), int(c["B"]), float(c["T"]) / int(c["N"])
    if str(c["kind"]) == "merton":
        dW, J = noise_streams.reg_trajectory_noise(int(c["seed"]), n, N, B, dt, float(c["lam"]), float(c["muJ"]), float(c["sigmaJ"]))
        assert abs(dW.astype(np.float64).sum() - c["dW_checksum"]) < 1e-9
    else:
        dW, J = None, noise_streams.vg_trajectory_noise(int(c["seed"]) + 1, n, N, B, dt, float(c["theta"]), float(c["kappa"]), float(c["sigmaJ"]))
    assert abs(J.astype(np.float64).sum() - c["J_checksum"]) < 1e-9
    return dW, J


@pytest.mark.parametrize("kind", ("merton", "vg"))
def test_oracle_follows_the_reference_reg_trajectory(kind):
    """300 consecutive Adam steps of the reference's own SolverGlobalSumLocalReg (the headline scheme, 1000 paths per step) on
    injected increments: the oracle's loss at every step and U(0, x0) after every update against the reference's."""
    c = load_case(TRAJ_REG[kind])
    dW, J = reg_trajectory_inputs(c)
    om, layout = oracle_of(c)
    B, n = int(c["B"]), int(c["nsteps"])
    th = torch.tensor(c["theta0"].copy())
    opt = KerasAdam(layout.total, float(c["lr"]))
    x0 = torch.tensor([[0.0, float(c["x0"])]], dtype=torch.float32)
    worst_l = worst_y = 0.0
    for k in range(n):
        t = th.clone().requires_grad_(True)
        nz = {"J": torch.tensor(J[k])[..., None]}
        if dW is not None:
            nz["dW"] = torch.tensor(dW[k])[..., None]
        loss = pricing_loss(om, "SumLocalReg", layout, t, nz, B)
        loss.backward()
        worst_l = max(worst_l, abs(float(loss.detach()) - c["losses"][k]) / abs(c["losses"][k]))
        opt.step(th, t.grad)
        y0 = float(mlp_forward(th, layout, 0, x0)[0, 0])
        worst_y = max(worst_y, abs(y0 - float(c["Y0_after_step"][k])))
    print(f"oracle vs reference over {n} {kind} Reg steps: worst loss rel {worst_l:.1e}, worst |Y0 - Y0_ref| {worst_y:.1e}")
    assert worst_l <= 2e-5 and worst_y <= 5e-6
    assert abs(float(c["Y0_after_step"][-1]) - float(c["Y0_report"])) < 1e-6


TRAJ_MFG = os.path.join(GOLD, "traj", "mfg_Global_200steps.npz")


def test_oracle_follows_the_reference_mfg_trajectory():
    """200 consecutive Adam steps of the reference's own MFG SolverGlobalFBSDE (couplage ON, both networks and both initial values
    trained together) on the increments it drew: the oracle's loss at every step and (Y0_hat, Y0) after every update."""
    c = load_case(TRAJ_MFG)
    om, layout = oracle_of(c)
    B, n = int(c["B"]), int(c["nsteps"])
    th = torch.tensor(c["theta0"].copy())
    opt = KerasAdam(layout.total, float(c["lr"]))
    worst_l = worst_y = 0.0
    for k in range(n):
        t = th.clone().requires_grad_(True)
        nz = {key: torch.tensor(c[key][k]) for key in ("dW0", "dW", "dN")}
        lh, li = mfg_loss(om, "Global", layout, t, nz, B)
        loss = lh + li
        loss.backward()
        worst_l = max(worst_l, abs(float(loss.detach()) - c["losses"][k]) / abs(c["losses"][k]))
        opt.step(th, t.grad)
        y = th[layout.y0_offset:layout.y0_offset + 2].numpy()
        worst_y = max(worst_y, float(np.abs(y - c["Y0_after_step"][k]).max()))
    print(f"oracle vs reference over {n} MFG Global steps: worst loss rel {worst_l:.1e}, worst |Y0 - Y0_ref| {worst_y:.1e}")
    assert worst_l <= 1e-5 and worst_y <= 5e-6


TRAJ_JUMP = {k: os.path.join(GOLD, "traj", f"{k}_Global_defaults_100steps.npz") for k in ("merton", "vg")}


def jump_trajectory_inputs(c):
    """The seeded increments of a default-shape Global trajectory (golden/noise_streams.py), checked against the fixture's checksums."""
    sys.path.insert(0, GOLD)
    import noise_streams
    kind = str(c["kind"])
    keys = ("lam", "muJ", "sigmaJ") if kind == "merton" else ("theta", "kappa", "sigmaJ")
    dW, J, JMC = noise_streams.jump_trajectory_noise(kind, int(c["seed"]), int(c["nsteps"]), int(c["N"]), int(c["B"]), int(c["M"]),
                                                     float(c["T"]) / int(c["N"]), {k: float(c[k]) for k in keys})
    assert abs(J.astype(np.float64).sum() - c["J_checksum"]) < 1e-9 and abs(JMC.astype(np.float64).sum() - c["JMC_checksum"]) < 1e-6
    if dW is not None:
        assert abs(dW.astype(np.float64).sum() - c["dW_checksum"]) < 1e-9
    return dW, J, JMC


@pytest.mark.parametrize("kind", ("merton", "vg"))
def test_oracle_follows_the_reference_default_shape_trajectory(kind):
    """The first 25 of 100 consecutive Adam steps of the reference's own SolverGlobalFBSDE at the default shapes of mainMerton.py /
    mainVG.py (10 paths, N = 50 / 30, 5000 compensator samples per time step): loss at every step and the trainable Y0 after every
    update.  (The GPU test follows all 100 steps.)"""
    c = load_case(TRAJ_JUMP[kind])
    dW, J, JMC = jump_trajectory_inputs(c)
    om, layout = oracle_of(c)
    B, n = int(c["B"]), 25
    th = torch.tensor(c["theta0"].copy())
    opt = KerasAdam(layout.total, float(c["lr"]))
    worst_l = worst_y = 0.0
    for k in range(n):
        t = th.clone().requires_grad_(True)
        nz = {"J": torch.tensor(J[k])[..., None], "JMC": torch.tensor(JMC[k])[..., None]}
        if dW is not None:
            nz["dW"] = torch.tensor(dW[k])[..., None]
        loss = pricing_loss(om, "Global", layout, t, nz, B)
        loss.backward()
        worst_l = max(worst_l, abs(float(loss.detach()) - c["losses"][k]) / abs(c["losses"][k]))
        opt.step(th, t.grad)
        worst_y = max(worst_y, abs(float(th[layout.y0_offset]) - float(c["Y0_after_step"][k])))
    print(f"oracle vs reference over {n} {kind} Global steps at the default shapes: worst loss rel {worst_l:.1e}, worst |Y0 - Y0_ref| {worst_y:.1e}")
    assert worst_l <= 3e-5 and worst_y <= 5e-6


def test_merton_closed_form_known_answers():
    om = MertonOracle(aLin=0.1, limit=30, d=1, dtype=torch.float64, **H.MERTON)
    assert abs(float(om.A(0, om.init(1))[0]) - 0.2714569268) < 1e-9
    x = torch.tensor([[0.8], [1.0], [1.2]], dtype=torch.float64)
    np.testing.assert_allclose(om.A(25, x).numpy(), [0.07911842, 0.20222704, 0.36822489], atol=2e-8)
    np.testing.assert_allclose(om.A(49, x).numpy(), [0.00217845, 0.10377461, 0.30217749], atol=2e-8)
    om10 = MertonOracle(aLin=0.1, limit=100, d=10, dtype=torch.float64, **H.MERTON)
    assert abs(float(om10.A(0, om10.init(1))[0]) - 0.1109224) < 2e-7
    om32 = MertonOracle(aLin=0.1, limit=30, d=1, **H.MERTON)
    assert abs(float(om32.A(0, om32.init(1))[0]) - 0.2714569268) < 2e-6


def test_vg_fft_known_answers():
    ov = VGOracle(aLin=0.1, dtype=torch.float64, **H.VG)
    assert abs(ov.correction - (-0.0796816965)) < 1e-9
    assert abs(float(ov.A(0, ov.init(1))[0]) - 0.1331402194) < 1e-8
    x = torch.tensor([[0.9], [1.0], [1.1]], dtype=torch.float64)
    np.testing.assert_allclose(ov.A(15, x).numpy(), [0.02925149, 0.08252713, 0.16135454], atol=2e-8)


def test_vg_direct_fourier_inversion_cross_checks_the_fft_price():
    """pricingModels.py:73-126 (VGmodelinvfourier.A, 1000-point trapezoid of the two Fourier integrals) against the Lewis/FFT
    price of pricingModels.py:156-179 (the oracle's VG.A): two independent pricers of the same model."""
    from deepfbsdejsolvers_b200.coupledPricing import VGmodelinvfourier, AbsCoupling
    ov = VGOracle(aLin=0.1, dtype=torch.float64, **H.VG)
    mv = VGmodelinvfourier(H.VG["T"], H.VG["N"], H.VG["r"], H.VG["theta"], H.VG["kappa"], H.VG["sigmaJ"], H.VG["K"], H.VG["x0"],
                           AbsCoupling(0.1))
    assert abs(float(mv.A(0, torch.ones(1))[0]) - 0.1331406) < 2e-7
    for i in (0, 7, 15, 29):
        x = torch.tensor([0.8, 0.9, 1.0, 1.1, 1.25], dtype=torch.float64)
        np.testing.assert_allclose(mv.A(i, x).numpy(), ov.A(i, x.reshape(-1, 1)).numpy().reshape(-1), atol=5e-5)   # du = 5 trapezoid


def test_exact_solution_makes_coupling_vanish():
    """with Y == A the coupling term is zero, so X follows the uncoupled jump-diffusion."""
    om = MertonOracle(aLin=0.1, limit=30, d=1, dtype=torch.float64, **H.MERTON)
    X = torch.tensor([[0.95], [1.3]], dtype=torch.float64)
    dW, J = torch.tensor([[0.02], [-0.1]], dtype=torch.float64), torch.tensor([[0.0], [0.15]], dtype=torch.float64)
    X1 = om.one_step(7, X, dW, J, om.A(7, X))
    np.testing.assert_allclose(X1.numpy(), (X * torch.exp(om.drift() * om.dt + om.sig * dW + J)).numpy(), rtol=1e-14)


def test_keras_adam_differs_from_torch_adam():
    th = torch.ones(4)
    g = torch.tensor([1e-8, 1e-3, 0.5, -2.0])
    opt = KerasAdam(4, 1e-3)
    opt.step(th, g)
    # first Keras step: lr*sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + 1e-7)
    expect = 1 - 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g.numpy()) / (np.sqrt(0.001 * g.numpy() ** 2) + 1e-7)
    np.testing.assert_allclose(th.numpy(), expect, rtol=1e-6)

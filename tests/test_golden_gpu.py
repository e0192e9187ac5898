"""GPU: the CUDA path against the outputs of the reference's own source files (tests/golden/*.npz, produced by
tests/golden/make_golden.py through the TensorFlow stand-in): one training step of every solver class - loss, tape
gradient, parameters after the Keras-form Adam update, reported Y0 - on the increments the reference drew
(compensator: the reference's hard-coded 5000 samples)."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from test_oracle_golden import load_case, CASES, OFF_CASES, TRAJ, off_masks

pytestmark = pytest.mark.gpu


def planes(a):
    """[N, n] (reference, d = 1) -> [N, 1, n]."""
    return np.ascontiguousarray(a[:, None, :])


MFG_KEYS = ("T", "R0", "jumpFactor", "alpha", "beta", "coeffOU", "A", "K", "pi", "p0", "p1", "f0", "f1", "theta", "C", "S0", "h1",
            "h2", "sig0", "sig", "alphaTarget", "coeffEqui")


@pytest.mark.parametrize("tensor_cores", (False, True), ids=("ffma", "tcgen05"))
@pytest.mark.parametrize("path", OFF_CASES, ids=[os.path.basename(p)[:-4] for p in OFF_CASES])
def test_cuda_reproduces_reference_couplage_off(ctx, path, tensor_cores):
    """couplage = 'OFF' two-phase training (MFGSolvers.py:92-115) against the reference's own outputs: objective weights (1, 0)
    then (0, 1), masked Keras-form Adam with the step counter carried over."""
    c = load_case(path)
    scheme, B = str(c["scheme"]), int(c["B"])
    layout = H.mfg_layout(scheme)
    s = H.native_mfg(ctx, dict(QAver=c["QAver"], jumpModel="stochastic", **{k: c[k] for k in MFG_KEYS}), scheme, layout,
                     tensor_cores=tensor_cores)
    s.set_theta(c["theta0"])
    s.reset_optimizer()
    hat, ind = off_masks(layout, scheme)
    for ph, mask, w in ((1, hat, (1.0, 0.0)), (2, ind, (0.0, 1.0))):
        s.set_weights(*w)
        s.set_noise(B, c[f"p{ph}_dW0"], c[f"p{ph}_dW"], c[f"p{ph}_dN"])
        out = s.grad(B)
        ref = c[f"p{ph}_loss"]
        assert abs(out[0] - ref) <= 2e-5 * abs(ref), (ph, out[0], ref)
        g_ref = c[f"p{ph}_grad"].astype(np.float64)
        err = np.abs(out[4:] * mask - g_ref).max() / np.abs(g_ref).max()
        assert err < (5e-5 if tensor_cores else 2e-5), f"phase {ph} gradient: rel-to-max error {err:.2e}"
        s.adam_step(float(c["lr"]), mask=ctx.to_device(mask))
    th = s.get_theta()
    solid = (np.abs(c["p1_grad"]) > 1e-3 * np.abs(c["p1_grad"]).max()) | (np.abs(c["p2_grad"]) > 1e-3 * np.abs(c["p2_grad"]).max())
    np.testing.assert_allclose(th[solid], c["theta2"][solid], rtol=0, atol=float(c["lr"]) * 2e-2)
    frozen = (hat + ind) == 0
    assert np.array_equal(th[frozen], c["theta0"][frozen])


@pytest.mark.parametrize("tensor_cores", (False, True), ids=("ffma", "tcgen05"))
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_cuda_reproduces_reference_step(ctx, path, tensor_cores):
    c = load_case(path)
    kind, scheme, B = str(c["kind"]), str(c["scheme"]), int(c["B"])
    if kind == "mfg":
        keys = ("T", "R0", "jumpFactor", "alpha", "beta", "coeffOU", "A", "K", "pi", "p0", "p1", "f0", "f1", "theta", "C", "S0", "h1",
                "h2", "sig0", "sig", "alphaTarget", "coeffEqui")
        layout = H.mfg_layout(scheme)
        s = H.native_mfg(ctx, dict(QAver=c["QAver"], jumpModel="stochastic", **{k: c[k] for k in keys}), scheme, layout,
                         tensor_cores=tensor_cores)
        s.set_theta(c["theta0"])
        s.set_noise(B, c["dW0"], c["dW"], c["dN"])
    else:
        layout = H.pricing_layout(kind, scheme, 1)
        par = {k: c[k] for k in (("T", "r", "muJ", "sigmaJ", "sigma", "lam", "K", "x0") if kind == "merton" else
                                 ("T", "r", "theta", "kappa", "sigmaJ", "K", "x0"))}
        par["N"] = int(c["N"])
        M = 0 if scheme.endswith("Reg") else c["JMC"].shape[1]
        s = H.native_pricing(ctx, kind, par, scheme, layout, d=1, M=M, tensor_cores=tensor_cores)
        s.set_theta(c["theta0"])
        s.set_noise(B, planes(c["dW"]) if "dW" in c else None, planes(c["J"]), planes(c["JMC"]) if "JMC" in c else None)
    out = s.grad(B)
    assert abs(out[0] - c["loss"]) <= 2e-5 * abs(c["loss"]), (out[0], c["loss"])
    g_ref = c["grad"].astype(np.float64)
    err = np.abs(out[4:] - g_ref).max() / np.abs(g_ref).max()
    print(f"{kind} {scheme} {'tcgen05' if tensor_cores else 'ffma'}: loss rel {abs(out[0] - c['loss']) / abs(c['loss']):.1e}, "
          f"gradient {err:.1e} of max")
    # tcgen05: 5e-5 for the compensator-free and MFG kernels; the jump schemes carry the cancellation between the own-jump row and
    # the compensator rows through bf16x3 products (tests/test_tc_gpu.py: JUMP_TC_GRAD_TOL)
    jump = kind != "mfg" and not scheme.endswith("Reg")
    assert err < ((3e-4 if jump else 5e-5) if tensor_cores else 2e-5), f"gradient: rel-to-max error {err:.2e}"
    s.adam_step(float(c["lr"]))
    th = s.get_theta()
    solid = np.abs(g_ref) > 1e-3 * np.abs(g_ref).max()
    np.testing.assert_allclose(th[solid], c["theta1"][solid], rtol=0, atol=float(c["lr"]) * 2e-2)
    # reported Y0 with the reference's post-update parameters
    s.set_theta(c["theta1"])
    if kind == "mfg":
        if scheme == "Global":
            y0h, y0 = c["theta1"][s.y0_offset], c["theta1"][s.y0_offset + 1]
        else:
            q0 = float(c["QAver"][0])
            y0h = s.net_forward(0, np.array([[0.0, q0, c["S0"], c["R0"]]], dtype=np.float32))[0, 0]
            y0 = s.net_forward(1, np.array([[0.0, q0, c["S0"], q0, c["S0"], c["R0"]]], dtype=np.float32))[0, 0]
        assert abs(y0h - c["Y0_hat_report"]) < 2e-6 and abs(y0 - c["Y0_report"]) < 2e-6
    elif scheme != "Global":
        y0 = s.net_forward(0, np.array([[0.0, c["x0"]]], dtype=np.float32))[0, 0]
        assert abs(y0 - c["Y0_report"]) < 2e-6


@pytest.mark.parametrize("tensor_cores", (False, True), ids=("ffma", "tcgen05"))
def test_cuda_follows_the_reference_training_trajectory(ctx, tensor_cores):
    """25 consecutive training steps of the reference's own SolverGlobalFBSDE (Merton, 5000 compensator samples per step) on the
    increments it drew: the CUDA path's loss at every step and its trainable Y0 after every Keras-form Adam update against the
    reference's - "the reference's own price" on the way (north star), step by step."""
    c = load_case(TRAJ)
    B, n, N = int(c["B"]), int(c["nsteps"]), int(c["N"])
    layout = H.pricing_layout("merton", "Global", 1)
    par = {k: c[k] for k in ("T", "r", "muJ", "sigmaJ", "sigma", "lam", "K", "x0")}
    par["N"] = N
    s = H.native_pricing(ctx, "merton", par, "Global", layout, d=1, M=c["JMC"].shape[2], tensor_cores=tensor_cores)
    s.set_theta(c["theta0"])
    s.reset_optimizer()
    worst_l = worst_y = 0.0
    for k in range(n):
        s.set_noise(B, planes(c["dW"][k]), planes(c["J"][k]), planes(c["JMC"][k]))
        out = s.grad(B)
        worst_l = max(worst_l, abs(out[0] - c["losses"][k]) / abs(c["losses"][k]))
        s.adam_step(float(c["lr"]))
        y0 = float(s.get_theta()[s.y0_offset])
        worst_y = max(worst_y, abs(y0 - float(c["Y0_after_step"][k])))
    print(f"trajectory ({'tcgen05' if tensor_cores else 'ffma'}): worst loss rel {worst_l:.1e}, worst |Y0 - Y0_ref| {worst_y:.1e} over {n} steps")
    assert worst_l <= 3e-5 and worst_y <= 5e-6
    th = s.get_theta()
    solid = np.abs(c["theta_final"] - c["theta0"]) > 0.2 * n * float(c["lr"])
    np.testing.assert_allclose(th[solid], c["theta_final"][solid], rtol=0, atol=0.03 * n * float(c["lr"]))


@pytest.mark.parametrize("kind", ("merton", "vg"))
@pytest.mark.parametrize("tensor_cores", (False, True), ids=("ffma", "tcgen05"))
def test_cuda_follows_the_reference_reg_trajectory(ctx, tensor_cores, kind):
    """300 consecutive training steps of the reference's own SolverGlobalSumLocalReg - the headline scheme, 1000 paths per step as its
    train() draws them - on injected increments (seeded stream, golden/noise_streams.py): the CUDA path's loss at every step and its
    U(0, x0) after every Keras-form Adam update against the reference's, and the final parameters."""
    from test_oracle_golden import TRAJ_REG, reg_trajectory_inputs
    c = load_case(TRAJ_REG[kind])
    dW, J = reg_trajectory_inputs(c)
    B, n, N = int(c["B"]), int(c["nsteps"]), int(c["N"])
    layout = H.pricing_layout(kind, "SumLocalReg", 1)
    keys = ("T", "r", "muJ", "sigmaJ", "sigma", "lam", "K", "x0") if kind == "merton" else ("T", "r", "theta", "kappa", "sigmaJ", "K", "x0")
    par = {k: c[k] for k in keys}
    par["N"] = N
    s = H.native_pricing(ctx, kind, par, "SumLocalReg", layout, d=1, M=0, tensor_cores=tensor_cores)
    s.set_theta(c["theta0"])
    s.reset_optimizer()
    x0 = np.array([[0.0, c["x0"]]], dtype=np.float32)
    worst_l = worst_y = 0.0
    for k in range(n):
        s.set_noise(B, None if dW is None else planes(dW[k]), planes(J[k]))
        out = s.grad(B)
        worst_l = max(worst_l, abs(out[0] - c["losses"][k]) / abs(c["losses"][k]))
        s.adam_step(float(c["lr"]))
        worst_y = max(worst_y, abs(float(s.net_forward(0, x0)[0, 0]) - float(c["Y0_after_step"][k])))
    th = s.get_theta()
    dth = float(np.abs(th - c["theta_final"]).max())
    print(f"{kind} Reg trajectory ({'tcgen05' if tensor_cores else 'ffma'}): worst loss rel {worst_l:.1e}, worst |Y0 - Y0_ref| {worst_y:.1e}, "
          f"max |theta - theta_ref| {dth:.1e} over {n} steps")
    # measured: ffma 6.7e-7 / 2.7e-7 / 8.9e-8, tcgen05 1.8e-6 / 9.8e-7 / 5.7e-7
    assert worst_l <= 1e-5 and worst_y <= 5e-6 and dth <= 5e-6


@pytest.mark.parametrize("tensor_cores", (False, True), ids=("ffma", "tcgen05"))
def test_cuda_follows_the_reference_mfg_trajectory(ctx, tensor_cores):
    """200 consecutive training steps of the reference's own MFG SolverGlobalFBSDE (couplage ON; Net_hat, Net and both trainable
    initial values in one Adam update) on the increments it drew: loss at every step, (Y0_hat, Y0) after every update, and the
    final parameters."""
    from test_oracle_golden import TRAJ_MFG
    c = load_case(TRAJ_MFG)
    B, n = int(c["B"]), int(c["nsteps"])
    layout = H.mfg_layout("Global")
    s = H.native_mfg(ctx, dict(QAver=c["QAver"], jumpModel="stochastic", **{k: c[k] for k in MFG_KEYS}), "Global", layout,
                     tensor_cores=tensor_cores)
    s.set_theta(c["theta0"])
    s.reset_optimizer()
    worst_l = worst_y = 0.0
    for k in range(n):
        s.set_noise(B, c["dW0"][k], c["dW"][k], c["dN"][k])
        out = s.grad(B)
        worst_l = max(worst_l, abs(out[0] - c["losses"][k]) / abs(c["losses"][k]))
        s.adam_step(float(c["lr"]))
        y = s.get_theta()[layout.y0_offset:layout.y0_offset + 2]
        worst_y = max(worst_y, float(np.abs(y - c["Y0_after_step"][k]).max()))
    th = s.get_theta()
    dth = float(np.abs(th - c["theta_final"]).max())
    print(f"MFG trajectory ({'tcgen05' if tensor_cores else 'ffma'}): worst loss rel {worst_l:.1e}, worst |Y0 - Y0_ref| {worst_y:.1e}, "
          f"max |theta - theta_ref| {dth:.1e} over {n} steps")
    # measured: ffma 2.1e-7 / 1.2e-7 / 2.4e-7, tcgen05 2.2e-7 / 1.2e-7 / 6.5e-6
    assert worst_l <= 5e-6 and worst_y <= 2e-6 and dth <= 3e-5


@pytest.mark.parametrize("kind", ("merton", "vg"))
@pytest.mark.parametrize("tensor_cores", (False, True), ids=("ffma", "tcgen05"))
def test_cuda_follows_the_reference_default_shape_trajectory(ctx, tensor_cores, kind):
    """100 consecutive training steps of the reference's own SolverGlobalFBSDE at the DEFAULT shapes of mainMerton.py / mainVG.py
    (BASELINE configs 1 and 2: 10 paths, N = 50 / 30, 5000 compensator samples redrawn at every time step) on seeded injected
    increments: the CUDA path (cluster of CTAs per path; jump rows on tcgen05 or FFMA) against the reference's loss at every step,
    its trainable Y0 after every update and its final parameters.

    With 10 paths per step Adam's normalisation moves every parameter by ~lr per step whatever the size of its gradient, so a
    gradient component below the rounding level of a path changes sign there and that parameter drifts away at up to lr per
    step.  On the fp32 FFMA kernels (rounding 1e-6 of the largest component) the loss stays within 2.5e-5 of the reference's
    over 100 steps; on the tcgen05 jump rows (bf16 hi+lo operands = 16 significant bits; measured 2e-5 / 2.2e-4 of the largest
    component for Merton / VG, whose 5000 samples are all non-zero and nearly cancel against the path's own jump -
    DESIGN 3.1b Numerics) it stays within 7e-5 (Merton) and 5e-3 (VG: 3e-4 after 10 steps).  The trainable Y0, whose gradient is
    large, follows to 2e-6 on every path.  `tensor_cores=False` selects the fp32 path."""
    from test_oracle_golden import TRAJ_JUMP, jump_trajectory_inputs
    c = load_case(TRAJ_JUMP[kind])
    dW, J, JMC = jump_trajectory_inputs(c)
    B, n, N = int(c["B"]), int(c["nsteps"]), int(c["N"])
    layout = H.pricing_layout(kind, "Global", 1)
    keys = ("T", "r", "muJ", "sigmaJ", "sigma", "lam", "K", "x0") if kind == "merton" else ("T", "r", "theta", "kappa", "sigmaJ", "K", "x0")
    par = {k: c[k] for k in keys}
    par["N"] = N
    s = H.native_pricing(ctx, kind, par, "Global", layout, d=1, M=int(c["M"]), tensor_cores=tensor_cores)
    s.set_theta(c["theta0"])
    s.reset_optimizer()
    worst_l = worst_y = 0.0
    marks = {}
    for k in range(n):
        s.set_noise(B, None if dW is None else planes(dW[k]), planes(J[k]), planes(JMC[k]))
        out = s.grad(B)
        worst_l = max(worst_l, abs(out[0] - c["losses"][k]) / abs(c["losses"][k]))
        s.adam_step(float(c["lr"]))
        worst_y = max(worst_y, abs(float(s.get_theta()[s.y0_offset]) - float(c["Y0_after_step"][k])))
        if k + 1 in (10, 25, 50):
            marks[k + 1] = (worst_l, worst_y)
    print("   running worst (loss rel, |Y0 - Y0_ref|):", {k: (f"{v[0]:.1e}", f"{v[1]:.1e}") for k, v in marks.items()})
    th = s.get_theta()
    dth = float(np.abs(th - c["theta_final"]).max())
    print(f"{kind} Global default-shape trajectory ({'tcgen05' if tensor_cores else 'ffma'}): worst loss rel {worst_l:.1e}, "
          f"worst |Y0 - Y0_ref| {worst_y:.1e}, max |theta - theta_ref| {dth:.1e} over {n} steps")
    # measured (loss rel at 10 / 100 steps; Y0): ffma merton 2.5e-5, 6e-8; ffma vg 1.5e-5, 0; tcgen05 merton 1.4e-6 / 7.0e-5, 6e-8;
    # tcgen05 vg 3.3e-4 / 5.0e-3, 2.3e-6
    tol10, tol100 = ((1e-5, 1e-4) if not tensor_cores else (1e-5, 3e-4) if kind == "merton" else (1e-3, 1.5e-2))
    assert marks[10][0] <= tol10 and worst_l <= tol100 and worst_y <= 1e-5
    if not tensor_cores:
        solid = np.abs(c["theta_final"] - c["theta0"]) > 0.2 * n * float(c["lr"])
        np.testing.assert_allclose(th[solid], c["theta_final"][solid], rtol=0, atol=0.03 * n * float(c["lr"]))


DIAG_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "diag", "*.npz")))


@pytest.mark.parametrize("path", DIAG_CASES, ids=[os.path.basename(p)[:-4] for p in DIAG_CASES])
def test_mfg_diagnostics_reproduce_the_reference(ctx, path):
    """simulateGlobalErr / followS of the drop-in MFG solver classes against the outputs of the reference's own
    methods (MFGSolvers.py:118-178, 436-459, executed through the TensorFlow stand-in) on the increments the reference drew."""
    from deepfbsdejsolvers_b200 import coupledMFG as cm
    c = load_case(path)
    assert len(DIAG_CASES) == 2
    scheme, nb = str(c["scheme"]), int(c["nb"])
    mm = cm.ModelCoupledFBSDE(QAver=c["QAver"], jumpModel="stochastic", **{k: c[k] for k in MFG_KEYS})
    wh, wi, method, cls = ((2, 3, "Global", "SolverGlobalFBSDE") if scheme == "Global" else (3, 4, "SumLocal", "SolverSumLocalFBSDE"))
    km = cm.kerasModels(cm.Net_hat, cm.Net, method, wh, wi, [20, 20], [22, 22], "tanh", "tanh")
    solver = getattr(cm, cls)(mm, km, 1e-3, "ON", ctx=ctx)
    solver.build().set_theta(c["theta0"])
    got = np.array(solver.simulateGlobalErr(nb, noise=(c["err_dW0"], c["err_dW"], c["err_dN"])))
    ref = c["err_result"]
    print(scheme, "simulateGlobalErr", got, "reference", ref)
    assert np.all(np.abs(got - ref) <= 2e-5 * np.abs(ref) + 1e-5), (got, ref)
    if "follow_result" in c:
        f = np.array(solver.followS(nb, noise=(c["follow_dW0"], c["follow_dW"], c["follow_dN"])))
        assert f.shape == c["follow_result"].shape
        assert np.abs(f - c["follow_result"]).max() <= 2e-6 + 2e-5 * np.abs(c["follow_result"]).max()

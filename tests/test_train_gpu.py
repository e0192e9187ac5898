"""GPU: the training call itself.  (1) fbsdej_solver_train_steps (CUDA graph: simulate -> forward -> adjoint -> fused
reduce + Adam + counters) must land on the same parameters as the same steps driven one C-ABI call at a time
(fbsdej_solver_grad_step -> fbsdej_adam_step -> fbsdej_bump_u32), for the fp32 FFMA kernels and for the tcgen05 kernels;
(2) the drop-in SolverGlobalSumLocalReg trains the Merton price towards the closed-form known answer 0.2714569
(pricingModels.py:40-49 at the mainMerton.py:57 parameters)."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme,d,M,tc", [("SumLocalReg", 10, 0, True), ("MultiStepReg", 1, 0, True), ("SumLocalReg", 1, 0, False),
                                            ("Global", 1, 48, False)])
def test_graph_replay_equals_stepwise_calls(ctx, scheme, d, M, tc):
    B, n, lr, seed = 300, 4, 3e-4, 99
    p = dict(H.MERTON, N=8)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 3)
    a = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=M, limit=30 if d == 1 else 100, tensor_cores=tc)
    b = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=M, limit=30 if d == 1 else 100, tensor_cores=tc)
    a.set_theta(theta); b.set_theta(theta)
    a.reset_optimizer(); b.reset_optimizer()
    losses = ctx.zeros(n)
    a.train_steps(seed, B, n, lr, loss_out=losses)
    step_loss = []
    for _ in range(n):
        out = b.grad_step(seed, B, B, 0)
        step_loss.append(float(ctx.to_host(out[:1]).numpy()[0]))
        b.adam_step(lr)
        b.bump_iteration()
    ctx.sync()
    ta, tb = a.get_theta(), b.get_theta()
    assert np.isfinite(ta).all() and np.abs(ta - theta).max() > 0
    assert np.abs(ta - tb).max() <= 1e-7 * max(1.0, np.abs(tb).max()), np.abs(ta - tb).max()
    la = ctx.to_host(losses).numpy()
    assert np.allclose(la, np.array(step_loss), rtol=1e-6, atol=0)
    assert int(ctx.to_host(a.t).numpy()[0]) == n and int(ctx.to_host(a.iteration).numpy()[0]) == int(ctx.to_host(b.iteration).numpy()[0])


def test_reg_solver_trains_to_closed_form(ctx):
    from deepfbsdejsolvers_b200 import coupledPricing as cp, set_seed
    set_seed(5)
    M = H.MERTON
    mm = cp.MertonJumpModel(M["T"], 20, M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(H.ALIN), 30)
    solver = cp.SolverGlobalSumLocalReg(mm, cp.Net(0, 1, [21, 21], "tanh"), cp.Net(0, 1, [21, 21], "tanh"), 2e-3, ctx=ctx)
    listY0, _ = solver.train(4, 10, 250, 16)         # train batch 1000 * 4 (SolversJumpDiff.py:435), val 100 * 10
    closed = float(mm.A(0, mm.init(1)).numpy()[0])
    assert abs(closed - 0.2714569) < 2e-5
    assert solver.lossList[-1] < solver.lossList[0]
    assert abs(float(listY0[-1]) - closed) < 0.02, (listY0, closed)


def test_checkpoint_resume_is_bit_exact(ctx, tmp_path):
    """params | m | v | t | iteration on disk; counter-based increments make the resumed run identical to the uninterrupted one."""
    B, lr, seed = 300, 3e-4, 7
    p = dict(H.MERTON, N=8)
    layout = H.pricing_layout("merton", "SumLocalReg", 10)
    theta = H.random_theta(layout, 4)
    a = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=10, limit=100, tensor_cores=True)
    a.set_theta(theta); a.reset_optimizer()
    a.train_steps(seed, B, 3, lr)
    ctx.sync()
    ck = str(tmp_path / "ck.npz")
    a.save_checkpoint(ck)
    a.train_steps(seed, B, 2, lr)
    ctx.sync()
    b = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=10, limit=100, tensor_cores=True)
    b.load_checkpoint(ck)
    b.train_steps(seed, B, 2, lr)
    ctx.sync()
    assert np.array_equal(a.get_theta(), b.get_theta())
    c = H.native_pricing(ctx, "merton", p, "MultiStepReg", layout, d=10, limit=100, tensor_cores=True)
    with pytest.raises(Exception):
        c.load_checkpoint(ck)


def test_driver_scripts_run(ctx, tmp_path, monkeypatch):
    """The mainMerton / mainVG / mainMFGComparison command lines on a tiny budget."""
    monkeypatch.chdir(tmp_path)
    from deepfbsdejsolvers_b200.coupledPricing import mainMerton, mainVG
    from deepfbsdejsolvers_b200.coupledMFG import mainMFGComparison
    cols = mainMerton.main(["--nEpochExt", "2", "--nEpoch", "3", "--batchSize", "4", "--methods", "Global,SumLocal1,SumMultiStepReg"])
    assert abs(cols["Y0_closed_formula"][0] - 0.2714569) < 2e-5 and len(cols["Y0_Global"]) == 2
    cols = mainVG.main(["--nEpochExt", "2", "--nEpoch", "3", "--batchSize", "4", "--methods", "Global,SumLocalReg"])
    assert abs(cols["Y0_closed_formula"][0] - 0.1331402) < 2e-5 and np.isfinite(cols["loss_Global"]).all()
    hY0, Y0 = mainMFGComparison.main(["--nEpochExt", "2", "--nEpoch", "3", "--batchSize", "16", "--methods", "Global,SumLocalReg"])
    assert len(hY0) == 2 and len(Y0[0]) == 2 and np.isfinite(np.array(Y0)).all()
    assert (tmp_path / "merton_Y0.csv").exists() and (tmp_path / "Y0List.csv").exists()
    # the network-shape flags of the reference's command lines (mainMerton.py:13-14, mainMFGComparison.py:14-17)
    cols = mainMerton.main(["--nEpochExt", "1", "--nEpoch", "2", "--batchSize", "4", "--nbLayer", "3", "--nbNeuron", "16",
                            "--methods", "Global,SumLocalReg"])
    assert np.isfinite(cols["Y0_Global"]).all() and np.isfinite(cols["Y0_SumLocalReg"]).all()
    hY0, Y0 = mainMFGComparison.main(["--nEpochExt", "1", "--nEpoch", "2", "--batchSize", "16", "--nbLayer_hat", "1", "--nbNeuron", "30",
                                      "--methods", "SumLocal"])
    assert np.isfinite(np.array(Y0)).all()


@pytest.mark.parametrize("name", ["SolverGlobalFBSDE", "SolverSumLocalFBSDE"])
def test_mfg_diagnostics(ctx, name):
    """simulateGlobalErr / followS (MFGSolvers.py:118-178) are reductions of the forward kernel's trajectory
    dump (itself parity-tested in test_parity_gpu.py::test_mfg): shapes, the initial state, and the identity
    cost = dt C sum_i mean(S_i) + h1 + h2 mean(S_N) between the two diagnostics (independent draws, Monte-Carlo tolerance)."""
    from deepfbsdejsolvers_b200 import coupledMFG as cm, set_seed
    set_seed(2)
    P = H.mfg_params(1)
    mm = cm.ModelCoupledFBSDE(**P)
    wh, wi, method = ((2, 3, "Global") if name == "SolverGlobalFBSDE" else (3, 4, "SumLocal"))
    km = cm.kerasModels(cm.Net_hat, cm.Net, method, wh, wi, [20, 20], [22, 22], "tanh", "tanh")
    solver = getattr(cm, name)(mm, km, 1e-3, "ON", ctx=ctx)
    solver.train(64, 128, 5, 1)
    nb = 40000
    c_hat, c_ind, mismatch = solver.simulateGlobalErr(nb)
    ah, sh, ai, si = solver.followS(nb)
    N = solver.native.N
    assert len(ah) == N + 1 and len(si) == N + 1 and np.isfinite([c_hat, c_ind, mismatch]).all() and mismatch >= 0
    assert abs(ah[0] - P["S0"]) < 1e-7 and sh[0] < 1e-7
    dt = P["T"] / N
    for cost, aver in ((c_hat, ah), (c_ind, ai)):
        ident = dt * P["C"] * float(np.sum(aver[:N])) + P["h1"] + P["h2"] * float(aver[N])
        assert abs(cost - ident) <= 0.02 * max(1.0, abs(ident)), (cost, ident)


@pytest.mark.parametrize("scheme", ["SumLocalReg", "MultiStepReg"])
def test_in_kernel_increments_equal_simulated_ones(ctx, scheme):
    """The tcgen05 Merton solvers draw their increments inside the forward sweep (no simulation kernel, nothing in HBM).
    Same Philox counters and arithmetic as fbsdej_solver_simulate, so the step on fused increments must reproduce
    simulate(seed, iteration) -> grad, for a shard with a path offset as well."""
    d, B, seed, it = 10, 333, 4242, 5
    p = dict(H.MERTON, N=7)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 8)
    a = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, limit=100, tensor_cores=True)
    b = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, limit=100, tensor_cores=True)
    a.set_theta(theta); b.set_theta(theta)
    for _ in range(it):
        b.bump_iteration()
    for off, cnt in ((0, B), (100, 200)):
        a.simulate(seed, it, cnt, path_offset=off)
        ref = a.grad(cnt, B_global=B)
        out = ctx.to_host(b.grad_step(seed, cnt, B, off)).numpy()
        assert abs(out[0] - ref[0]) <= 1e-6 * abs(ref[0]), (out[0], ref[0])
        assert np.abs(out[4:] - ref[4:]).max() <= 1e-6 * np.abs(ref[4:]).max()


def test_fixed_trajectory_replay_matches_host_recursion(ctx):
    """MFGSolutionsFixedTrajectory on pre-drawn [nbSimul, N+1] increments.  The fused kernel's replay (states,
    controls) against the reference recursion stepped on the host (ModelCoupledFBSDE.oneStepFrom / calpha / calpha_hat,
    MFGModel.py:58-89) with the same networks."""
    import torch
    from deepfbsdejsolvers_b200 import coupledMFG as cm, set_seed
    set_seed(3)
    P = H.mfg_params(1)
    mm = cm.ModelCoupledFBSDE(**P)
    km = cm.kerasModels(cm.Net_hat, cm.Net, "SumLocalReg", 1, 1, [20, 20], [22, 22], "tanh", "tanh")
    solver = cm.SolverGlobalSumLocalReg(mm, km, 1e-3, "ON", ctx=ctx)
    solver.train(64, 128, 5, 1)
    nb, N = 48, mm.N
    rng = np.random.default_rng(0)
    dW0 = (np.sqrt(mm.dt) * rng.standard_normal((nb, N + 1))).astype(np.float32)
    dW = (np.sqrt(mm.dt) * rng.standard_normal((nb, N + 1))).astype(np.float32)
    dN = rng.poisson(0.3, (nb, N + 1)).astype(np.float32)
    sol = cm.MFGSolutionsFixedTrajectory(mm, km, "SumLocalReg", dW0, dW, dN, ctx=ctx)
    sol.simulateAllProcesses(nb)
    assert sol.hS.shape == (nb, N + 1) and sol.alpha.shape == (nb, N + 1) and sol.meanhQ.shape == (N + 1,)
    # host recursion with the same (trained) networks
    s = solver.native
    mm.init(nb)
    for i in range(N + 1):
        t = np.full(nb, i * mm.dt, dtype=np.float32)
        hY = s.net_forward(0, np.stack([t, mm.hQ.numpy(), mm.hS.numpy(), mm.R.numpy()], 1))[:, 0]
        Y = s.net_forward(1, np.stack([t, mm.Q.numpy(), mm.S.numpy(), mm.hQ.numpy(), mm.hS.numpy(), mm.R.numpy()], 1))[:, 0]
        hYt, Yt = torch.from_numpy(hY), torch.from_numpy(Y)
        for name, ref in (("hQ", mm.hQ), ("Q", mm.Q), ("R", mm.R), ("hS", mm.hS), ("S", mm.S)):
            got = getattr(sol, name)[:, i]
            assert np.abs(got - ref.numpy()).max() <= 2e-4 * max(1.0, np.abs(ref.numpy()).max()), (name, i)
        assert np.abs(sol.alpha_hat[:, i] - mm.calpha_hat(hYt).numpy()).max() <= 5e-4 * max(1.0, np.abs(sol.alpha_hat[:, i]).max())
        assert np.abs(sol.alpha[:, i] - mm.calpha(hYt, Yt).numpy()).max() <= 5e-4 * max(1.0, np.abs(sol.alpha[:, i]).max())
        if i < N:
            mm.oneStepFrom(torch.from_numpy(dW0[:, i]), torch.from_numpy(dW[:, i]), torch.from_numpy(dN[:, i]), hYt, Yt)
    mean_cost, std_cost = sol.objectiveFunction()
    assert np.isfinite([mean_cost, std_cost]).all() and sol.price(mm.pi, sol.alpha_hat).shape == (nb, N + 1)
    # the Global replay (BSDE-propagated Y) runs through the same path
    kg = cm.kerasModels(cm.Net_hat, cm.Net, "Global", 2, 3, [20, 20], [22, 22], "tanh", "tanh")
    cm.SolverGlobalFBSDE(mm, kg, 1e-3, "ON", ctx=ctx).train(32, 64, 2, 1)
    solg = cm.MFGSolutionsFixedTrajectory(mm, kg, "Global", dW0, dW, dN, ctx=ctx)
    solg.simulateAllProcesses(nb)
    assert np.isfinite(solg.objectiveFunction()).all()


def test_mfg_graph_replay_equals_stepwise_calls(ctx):
    """The MFG training call (CUDA graph: Cox-count simulation -> two-network forward -> adjoint -> fused reduce + Adam)
    against the same steps driven call by call."""
    B, n, lr, seed = 200, 3, 1e-3, 17
    P = H.mfg_params(1)
    layout = H.mfg_layout("Global")
    theta = H.random_theta(layout, 6)
    a, b = H.native_mfg(ctx, P, "Global", layout), H.native_mfg(ctx, P, "Global", layout)
    a.set_theta(theta); b.set_theta(theta)
    a.reset_optimizer(); b.reset_optimizer()
    a.train_steps(seed, B, n, lr)
    for _ in range(n):
        b.grad_step(seed, B, B, 0)
        b.adam_step(lr)
        b.bump_iteration()
    ctx.sync()
    ta, tb = a.get_theta(), b.get_theta()
    assert np.isfinite(ta).all() and np.abs(ta - theta).max() > 0
    assert np.abs(ta - tb).max() <= 1e-7 * max(1.0, np.abs(tb).max())


def _pricing_solver(ctx, seed=5):
    from deepfbsdejsolvers_b200 import coupledPricing as cp, set_seed
    set_seed(seed)
    M = H.MERTON
    mm = cp.MertonJumpModel(M["T"], 8, M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(H.ALIN), 30)
    s = cp.SolverGlobalSumLocalReg(mm, cp.Net(0, 1, [21, 21], "tanh"), cp.Net(0, 1, [21, 21], "tanh"), 1e-3, ctx=ctx)
    s.build()                  # (the networks draw their initial weights here: right after set_seed)
    return s


def _mfg_solver(ctx, couplage="ON", cls="SolverGlobalFBSDE", seed=6):
    from deepfbsdejsolvers_b200 import coupledMFG as cm, set_seed
    set_seed(seed)
    p = H.mfg_params(1)
    mm = cm.ModelCoupledFBSDE(**p)
    glob = cls == "SolverGlobalFBSDE"
    km = cm.kerasModels(cm.Net_hat, cm.Net, "Global" if glob else "SumLocalReg", 2 if glob else 1, 3 if glob else 1, [20, 20], [22, 22],
                        "tanh", "tanh")
    s = getattr(cm, cls)(mm, km, 1e-3, couplage, ctx=ctx)
    s.build()
    return s


@pytest.mark.parametrize("kind", ["pricing", "mfg"])
def test_class_level_save_load_train_equals_uninterrupted_train(ctx, tmp_path, kind):
    """Solver.save -> a new Solver.load -> train continues with the restored Adam m, v, t and Philox iteration: the same
    parameters, bit for bit, as one uninterrupted train() (the reference keeps nothing on disk).  A name without
    the .npz suffix works on both sides."""
    make = _pricing_solver if kind == "pricing" else _mfg_solver
    full, first, second = make(ctx), make(ctx), make(ctx)
    full.train(1 if kind == "pricing" else 32, 2, 3, 4)
    first.train(1 if kind == "pricing" else 32, 2, 3, 2)
    ck = str(tmp_path / "ck")                      # no suffix
    first.save(ck)
    second.load(ck)
    second.train(1 if kind == "pricing" else 32, 2, 3, 2)
    assert np.array_equal(full.native.get_theta(), second.native.get_theta())
    assert len(second.listY0) == 4 and np.array_equal(np.array(full.listY0), np.array(second.listY0))
    # a train() that does NOT follow a load() starts a fresh optimizer, as the reference does (SolversJumpDiff.py:55)
    second.train(1 if kind == "pricing" else 32, 2, 1, 1)
    assert int(ctx.to_host(second.native.t).numpy()[0]) == 1


@pytest.mark.parametrize("cls", ["SolverGlobalFBSDE", "SolverGlobalSumLocalReg"])
def test_mfg_couplage_off_class_training(ctx, cls):
    """couplage = 'OFF' through the drop-in class (MFGSolvers.py:92-115): phase 1 updates only model_hat's variables on the
    projected player's loss, phase 2 only model's on the individual player's loss, one Adam step counter over both phases -
    the same parameters as the two phases driven by hand through the C-ABI."""
    B, ne, next_ = 32, 3, 2
    s = _mfg_solver(ctx, "OFF", cls)
    s.build()
    theta0 = s.native.get_theta().copy()
    s.train(B, 8, ne, next_)
    th = s.native.get_theta()
    assert int(ctx.to_host(s.native.t).numpy()[0]) == 2 * ne * next_          # the counter carried over
    assert len(s.listY0_hat) == next_ and len(s.listY0) == next_
    ref = _mfg_solver(ctx, "OFF", cls)
    n = ref.build()
    assert np.array_equal(n.get_theta(), theta0)
    n.reset_optimizer()
    for w, which in (((1.0, 0.0), "hat"), ((0.0, 1.0), "ind")):
        n.set_weights(*w)
        n.train_steps(ref.seed, B, ne * next_, ref.lRate, mask=ref._mask(which))
        ctx.sync()
        if which == "hat":
            mid = n.get_theta()
            ind = ref._mask("ind").cpu().numpy() > 0
            assert np.array_equal(mid[ind], theta0[ind]) and np.abs(mid - theta0).max() > 0
    assert np.array_equal(n.get_theta(), th)
    hat = ref._mask("hat").cpu().numpy() > 0
    assert np.array_equal(th[hat], mid[hat])                                   # phase 2 left model_hat alone


def test_tensor_cores_true_on_an_unsupported_shape_raises(ctx):
    from deepfbsdejsolvers_b200 import coupledPricing as cp, coupledMFG as cm
    M = H.MERTON
    mm = cp.MertonJumpModel(M["T"], 8, M["r"], M["muJ"], M["sigmaJ"], M["sigma"], M["lam"], M["K"], M["x0"], cp.AbsCoupling(H.ALIN), 30)
    with pytest.raises(ValueError):
        cp.SolverGlobalSumLocalReg(mm, cp.Net(0, 1, [40, 40], "tanh"), cp.Net(0, 1, [40, 40], "tanh"), 1e-3, ctx=ctx, tensor_cores=True).build()
    mfg = cm.ModelCoupledFBSDE(**H.mfg_params(1))
    km = cm.kerasModels(cm.Net_hat, cm.Net, "SumLocalReg", 1, 1, [30, 30], [30, 30], "tanh", "tanh")
    with pytest.raises(ValueError):
        cm.SolverGlobalSumLocalReg(mfg, km, 1e-3, "ON", ctx=ctx, tensor_cores=True).build()

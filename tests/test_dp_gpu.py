"""GPU: data-parallel invariants on ONE device (SURVEY section 4 iv).  Philox counter word 0 is the global path id, so
G virtual shards (path_offset = shard start, B_global weight) must reproduce the single-shard noise bit for bit and,
summed, the single-shard loss / gradient up to fp32 summation order."""
import numpy as np
import pytest

import helpers as H
from deepfbsdejsolvers_b200.solver_base import shard

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme,M", [("SumLocalReg", 0), ("Global", 64)])
def test_virtual_shards_reproduce_single_gpu(ctx, scheme, M):
    d, B, seed = 10, 1000, 1234
    p = dict(H.MERTON, N=10)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 5)
    full = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=M, limit=100)
    full.set_theta(theta)
    full.simulate(seed, 3, B)
    ref = full.grad(B)
    pa, pj, *_ = full.get_noise()
    dW_full = full.read_device(pa, p["N"] * d * B).reshape(p["N"], d, B)
    J_full = full.read_device(pj, p["N"] * d * B).reshape(p["N"], d, B)
    for world in (2, 4):
        acc = np.zeros_like(ref, dtype=np.float64)
        for rank in range(world):
            off, cnt = shard(B, rank, world)
            s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=M, limit=100)
            s.set_theta(theta)
            s.simulate(seed, 3, cnt, path_offset=off)
            qa, qj, *_ = s.get_noise()
            dW = s.read_device(qa, p["N"] * d * cnt).reshape(p["N"], d, cnt)
            J = s.read_device(qj, p["N"] * d * cnt).reshape(p["N"], d, cnt)
            assert np.array_equal(dW, dW_full[:, :, off:off + cnt]) and np.array_equal(J, J_full[:, :, off:off + cnt])
            acc += s.grad(cnt, B_global=B)
        assert abs(acc[0] - ref[0]) <= 2e-6 * abs(ref[0])
        scale = np.abs(ref[4:]).max()
        assert np.abs(acc[4:] - ref[4:]).max() <= 2e-5 * scale


def test_simulated_increment_statistics(ctx):
    """Moments of the simulated increments against the distributions the reference draws from
    (sqrt(dt) N(0,1); compound Poisson(lam dt) with N(muJ, sigJ^2) sizes; SolversJumpDiff.py:30-34, pricingModels.py:57-61)."""
    d, B, N = 10, 200000, 4
    p = dict(H.MERTON, N=N)
    layout = H.pricing_layout("merton", "SumLocalReg", d)
    s = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=d, limit=100)
    s.simulate(7, 0, B)
    pa, pj, *_ = s.get_noise()
    dW = s.read_device(pa, N * d * B).astype(np.float64)
    J = s.read_device(pj, N * d * B).astype(np.float64)
    dt, lam, sj = p["T"] / N, p["lam"], p["sigmaJ"]
    n = dW.size
    assert abs(dW.mean()) < 5 * np.sqrt(dt / n)
    assert abs(dW.var() / dt - 1) < 5 * np.sqrt(2.0 / n)
    assert abs((dW ** 4).mean() / dt ** 2 - 3) < 0.05
    p0 = np.exp(-lam * dt)
    assert abs((J == 0).mean() - p0) < 5 * np.sqrt(p0 * (1 - p0) / n)
    assert abs(J.var() - lam * dt * sj ** 2) < 0.02 * lam * dt * sj ** 2        # Var = E[dN] sigJ^2 (muJ = 0)
    s.simulate(7, 1, B)
    assert not np.array_equal(s.read_device(pa, 1000), dW[:1000].astype(np.float32))   # new iteration, new draws


def _two_ranks_through_the_fused_exchange(ctx, make, theta, B, seed, steps, lr, noisy_entries):
    """Two ranks of one process on one device (two streams): fbsdej_solver_train_steps_dp exchanges the [loss | gradient]
    vector through the peers' buffers inside the finishing kernel.  Both ranks must end with bit-identical parameters, equal to
    single-rank training on the whole batch up to the fp32 summation order of the gradient."""
    import threading
    from deepfbsdejsolvers_b200 import Context
    world = 2
    full = make(ctx)
    full.set_theta(theta)
    loss_full = ctx.zeros(steps)
    full.train_steps(seed, B, steps, lr, loss_out=loss_full)
    ctx.sync()
    ranks, losses = [], []
    for r in range(world):
        c = Context(ctx.index)
        s = make(c)
        s.set_theta(theta)
        off, cnt = shard(B, r, world)
        s.grad_step(seed, cnt, B, off)                   # sizes every buffer: no allocation while a peer's kernel waits
        c.sync()
        ranks.append((c, s, off, cnt))
        losses.append(c.zeros(steps))
    bufs = []
    for r, (c, s, off, cnt) in enumerate(ranks):
        s.dp_init(r, world)
        bufs.append(s.dp_buffer())
    for c, s, off, cnt in ranks:
        s.dp_connect(raw_ptrs=bufs)
    errs = []

    def run(r):
        c, s, off, cnt = ranks[r]
        try:
            s.train_steps_dp(seed, cnt, B, off, steps, lr, loss_out=losses[r])
            c.sync()
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errs and not any(t.is_alive() for t in th), errs
    t0 = ranks[0][0].to_host(ranks[0][1].theta).numpy()
    t1 = ranks[1][0].to_host(ranks[1][1].theta).numpy()
    assert np.array_equal(t0, t1)
    l0, l1 = ranks[0][0].to_host(losses[0]).numpy(), ranks[1][0].to_host(losses[1]).numpy()
    assert np.array_equal(l0, l1) and np.isfinite(l0).all()
    lf, tf = ctx.to_host(loss_full).numpy(), ctx.to_host(full.theta).numpy()
    assert np.abs(l0 - lf).max() <= 2e-5 * np.abs(lf).max(), (l0, lf)
    # Entries whose gradient is zero or changes sign from step to step (the output bias of the jump network is exactly zero in
    # exact arithmetic) follow the fp32 summation order: Adam turns their rounding noise into +-lr steps.  They are recognised by
    # how little they moved; every entry that moved steadily must agree, the others stay within 2 % of the distance lr * steps.
    # (measured, Global / 64 samples: the jump network's output bias drifts by 4e-5 - its gradient is rounding noise - and through
    # it the steady entries by up to 2.1e-6 = 3.5e-4 of the distance travelled; compensator-free solvers: nothing above 2e-4)
    dev, move = np.abs(t0 - tf), np.abs(tf - theta)
    far = dev > 2e-4 * move.max() + 1e-7
    if noisy_entries == 0:
        assert not far.any(), (far.sum(), dev.max())
    else:
        steady = move > 0.5 * steps * lr
        assert not (steady & (dev > 1e-3 * move.max() + 1e-7)).any(), (np.nonzero(steady & far)[0], dev[steady].max())
        assert dev.max() <= 0.02 * steps * lr, dev.max()
        assert (far & ~steady).sum() <= max(noisy_entries, int((~steady).sum())), (far.sum(), dev.max())


@pytest.mark.parametrize("scheme,M,d", [("SumLocalReg", 0, 10), ("Global", 64, 1)])
def test_exchange_inside_the_finishing_kernel(ctx, scheme, M, d):
    p = dict(H.MERTON, N=10)
    layout = H.pricing_layout("merton", scheme, d)
    kw = dict(d=d, M=M, limit=30 if d == 1 else 100, tensor_cores=True, price_table=d > 1)
    _two_ranks_through_the_fused_exchange(ctx, lambda c: H.native_pricing(c, "merton", p, scheme, layout, **kw),
                                          H.random_theta(layout, 5), 1000, 99, 6, 1e-3, 0 if M == 0 else 1)


def test_exchange_inside_the_finishing_kernel_mfg(ctx):
    p = H.mfg_params(1, "stochastic")
    layout = H.mfg_layout("MultiStep")
    _two_ranks_through_the_fused_exchange(ctx, lambda c: H.native_mfg(c, p, "MultiStep", layout, tensor_cores=True),
                                          H.random_theta(layout, 6), 256, 7, 5, 1e-3, 2)


def test_missing_peer_voids_the_step_and_raises(ctx, monkeypatch):
    """A rank whose peer never delivers its vector must neither hang nor poison its parameters: after FBSDEJ_DP_TIMEOUT_MS the
    step is void (parameters, Adam slots and counters untouched), the error word is raised in BOTH ranks' buffers and
    fbsdej_solver_dp_check fails on both."""
    from deepfbsdejsolvers_b200 import Context, FbsdejError
    monkeypatch.setenv("FBSDEJ_DP_TIMEOUT_MS", "200")
    p = dict(H.MERTON, N=6)
    layout = H.pricing_layout("merton", "SumLocalReg", 10)
    theta = H.random_theta(layout, 5)
    B, ranks, bufs = 400, [], []
    for r in range(2):
        c = Context(ctx.index)
        s = H.native_pricing(c, "merton", p, "SumLocalReg", layout, d=10, limit=100, tensor_cores=True)
        s.set_theta(theta); s.reset_optimizer()
        s.grad_step(3, B // 2, B, r * (B // 2))
        c.sync()
        s.dp_init(r, 2)
        bufs.append(s.dp_buffer())
        ranks.append((c, s))
    for c, s in ranks:
        s.dp_connect(raw_ptrs=bufs)
    c0, s0 = ranks[0]
    loss = c0.zeros(2)
    s0.train_steps_dp(3, B // 2, B, 0, 2, 1e-3, loss_out=loss)     # rank 1 never steps
    with pytest.raises(FbsdejError, match="timed out"):
        s0.dp_check()
    assert np.array_equal(s0.get_theta(), theta)                     # not updated, not NaN
    assert int(c0.to_host(s0.t).numpy()[0]) == 0 and not np.abs(c0.to_host(s0.m).numpy()).any()
    assert np.isnan(c0.to_host(loss).numpy()).all()                  # the record says the steps were void
    with pytest.raises(FbsdejError, match="timed out"):
        ranks[1][1].dp_check()                                       # the peer sees the failure too

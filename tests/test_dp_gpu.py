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

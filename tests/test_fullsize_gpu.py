"""GPU: BASELINE.json's full size (config 3: Merton d = 10, B = 2^16 paths x N = 100 steps, SolverGlobalSumLocalReg on the tcgen05
kernels), checked through size-independent properties - the oracle cannot run this size in seconds:
  * determinism: the same (seed, iteration) gives bit-identical losses, gradients and post-Adam parameters;
  * the batch mean is shard-additive: two virtual shards (path offsets 0 and B/2, weights 1/B) sum to the full-batch loss and
    gradient up to fp32 summation order - the data-parallel invariant at full size;
  * increments drawn inside the forward sweep == increments materialised by the simulation kernel;
  * the tcgen05 kernels agree with the fp32 FFMA kernels on the same increments (1e-5 loss, 2e-4 of the largest gradient);
  * a few hundred training steps reduce the validation loss."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

B, D, SEED = 1 << 16, 10, 2026
P = dict(H.MERTON, N=100)


def _solver(ctx, theta, tc=True):
    layout = H.pricing_layout("merton", "SumLocalReg", D)
    s = H.native_pricing(ctx, "merton", P, "SumLocalReg", layout, d=D, limit=100, tensor_cores=tc)
    s.set_theta(theta)
    s.reset_optimizer()
    return s


@pytest.fixture(scope="module")
def theta():
    return H.random_theta(H.pricing_layout("merton", "SumLocalReg", D), 77)


def test_determinism_and_shard_additivity(ctx, theta):
    a, b = _solver(ctx, theta), _solver(ctx, theta)
    full = ctx.to_host(a.grad_step(SEED, B, B, 0)).numpy().copy()
    again = ctx.to_host(b.grad_step(SEED, B, B, 0)).numpy().copy()
    assert np.array_equal(full, again)
    acc = np.zeros_like(full, dtype=np.float64)
    for off in (0, B // 2):
        acc += ctx.to_host(b.grad_step(SEED, B // 2, B, off)).numpy()
    # (the shards are tiled differently from the full batch, so the bf16x3 weight-gradient sums group differently)
    err_l, err_g = abs(acc[0] - full[0]) / abs(full[0]), np.abs(acc[4:] - full[4:]).max() / np.abs(full[4:]).max()
    print(f"shard additivity: loss {err_l:.1e}, gradient {err_g:.1e} of max")
    assert err_l <= 2e-6 and err_g <= 5e-5, (err_l, err_g)
    a.train_steps(SEED, B, 3, 3e-4); b.train_steps(SEED, B, 3, 3e-4)
    ctx.sync()
    assert np.array_equal(a.get_theta(), b.get_theta())


def test_fused_increments_and_ffma_agreement(ctx, theta):
    tc, ff = _solver(ctx, theta), _solver(ctx, theta, tc=False)
    fused = ctx.to_host(tc.grad_step(SEED, B, B, 0)).numpy().copy()      # increments drawn inside the forward sweep
    tc.simulate(SEED, 0, B)
    mat = tc.grad(B)                                                     # same counters, materialised by sim_merton_kernel
    assert abs(fused[0] - mat[0]) <= 1e-6 * abs(mat[0]) and np.abs(fused[4:] - mat[4:]).max() <= 1e-6 * np.abs(mat[4:]).max()
    ff.simulate(SEED, 0, B)
    ref = ff.grad(B)                                                     # fp32 FFMA kernels on the same increments
    assert abs(mat[0] - ref[0]) <= 1e-5 * abs(ref[0]), (mat[0], ref[0])
    assert np.abs(mat[4:] - ref[4:]).max() <= 2e-4 * np.abs(ref[4:]).max()


def test_training_reduces_the_validation_loss(ctx, theta):
    s = _solver(ctx, theta)
    s.simulate(SEED ^ 0x55, 0, B)
    before = float(s.loss(B)[0])
    s.train_steps(SEED, B, 300, 1e-3)
    s.simulate(SEED ^ 0x55, 0, B)
    after = float(s.loss(B)[0])
    assert np.isfinite(after) and after < 0.7 * before, (before, after)


@pytest.mark.parametrize("Bt", [56831, 56832, 56833, 60000, 75776, 75777, 113664, 131073])
def test_tile_map_edges(ctx, theta, Bt):
    """Batch sizes around the switches of the tile map (pricing.cuh: make_tile_map - uniform 128-row tiles below 3 warps per CTA
    slot, mixed 128 / 96-row tiles above, a second wave beyond 4 warps per slot): the tcgen05 kernels on in-kernel increments against
    the fp32 FFMA kernels on the same (materialised) increments, three time steps."""
    p = dict(P, N=3)
    layout = H.pricing_layout("merton", "SumLocalReg", D)
    tc = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=D, limit=100, tensor_cores=True)
    ff = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=D, limit=100, tensor_cores=False)
    tc.set_theta(theta); ff.set_theta(theta)
    fused = ctx.to_host(tc.grad_step(SEED, Bt, Bt, 0)).numpy().copy()
    ff.simulate(SEED, 0, Bt)
    ref = ff.grad(Bt)
    assert abs(fused[0] - ref[0]) <= 1e-5 * abs(ref[0]), (fused[0], ref[0])
    assert np.abs(fused[4:] - ref[4:]).max() <= 1e-4 * np.abs(ref[4:]).max()
    # and the trajectory dump finds every path again (untile through the same map)
    tc.simulate(SEED, 0, Bt)
    _, tx, _, _ = tc.loss(Bt, traj=True)
    _, fx, _, _ = ff.loss(Bt, traj=True)
    assert np.abs(tx - fx).max() <= 2e-6 + 1e-5 * np.abs(fx).max()

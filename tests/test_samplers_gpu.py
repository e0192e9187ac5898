"""GPU: the device samplers (csrc/sim_kernels.cu, sim_device.cuh) draw the distributions the reference draws from - checked through
fbsdej_solver_simulate + fbsdej_solver_get_noise, i.e. on exactly the increments the training path consumes:

  Merton  dW = sqrt(dt) N(0,1); J = dN muJ + sigJ sqrt(dN) eps, dN ~ Poisson(lam dt)      SolversJumpDiff.py:30-34, pricingModels.py:57-61
  VG      gamma ~ Gamma(shape dt/kappa, scale kappa); J = theta gamma + sigJ sqrt(gamma) eps     pricingModels.py:188-191
  MFG     dN_i ~ Poisson(lam(hQ_i) dt), lam = beta (e^{alpha hQ} - 1) or jumpFactor, intensity frozen at the step start;
          hQ_{i+1} = hQ_i + kappa (Qbar_{i+1} - hQ_i) dt + sig0 dW0_i                            MFGModel.py:47-54, 70

KS distances are compared with the 0.1 % critical value 1.95 / sqrt(n) (n ~ 10^5 .. 10^6; a wrong shape, scale or tail shows up
as a distance of 10^-2 or more); discrete laws with the total-variation distance to the SciPy pmf."""
import numpy as np
import pytest
from scipy import stats

import helpers as H

pytestmark = pytest.mark.gpu


def ks_distance(x, cdf):
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = x.size
    F = cdf(x)
    return max(np.abs(F - np.arange(1, n + 1) / n).max(), np.abs(F - np.arange(0, n) / n).max())


def _merton_increments(ctx, lam, N, d, B, seed, muJ=0.0):
    p = dict(H.MERTON, N=N, lam=lam, muJ=muJ)
    layout = H.pricing_layout("merton", "SumLocalReg", d)
    s = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=d, limit=100)
    s.simulate(seed, 0, B)
    pa, pj, *_ = s.get_noise()
    return p, s.read_device(pa, N * d * B).astype(np.float64), s.read_device(pj, N * d * B).astype(np.float64)


def test_merton_brownian_and_single_jump_distributions(ctx):
    """Reference defaults (lam dt = 0.06): Brownian increments KS vs N(0, dt); the jump sizes given one jump KS vs N(muJ, sigJ^2)
    (their quantile comes from the residual of the Poisson inversion - the sampler's own construction)."""
    N, d, B = 5, 10, 100000
    p, dW, J = _merton_increments(ctx, 3.0, N, d, B, seed=11, muJ=0.05)
    dt, lam, mu, sj = p["T"] / N, 3.0, 0.05, p["sigmaJ"]
    n = dW.size
    D = ks_distance(dW[:400000] / np.sqrt(dt), stats.norm.cdf)
    assert D < 1.95 / np.sqrt(400000), D
    nz = J[J != 0.0]
    p1 = 1.0 - np.exp(-lam * dt)
    assert abs(nz.size / n - p1) < 5 * np.sqrt(p1 * (1 - p1) / n)
    # mixture over the count: n jumps ~ N(n muJ, n sigJ^2), weights Poisson(n) / P(n >= 1)
    w = stats.poisson.pmf(np.arange(1, 12), lam * dt) / p1
    cdf = lambda x: sum(w[k - 1] * stats.norm.cdf(x, k * mu, sj * np.sqrt(k)) for k in range(1, 12))
    D = ks_distance(nz, cdf)
    print("Merton: jumps", nz.size, "KS", D)
    assert D < 1.95 / np.sqrt(nz.size), D
    # increments of different (path, asset, step) cells are uncorrelated
    a = dW.reshape(N, d, B)
    assert abs(np.corrcoef(a[0, 0], a[0, 1])[0, 1]) < 5 / np.sqrt(B) and abs(np.corrcoef(a[0, 0], a[1, 0])[0, 1]) < 5 / np.sqrt(B)
    assert abs(np.corrcoef(a[0, 0, :-1], a[0, 0, 1:])[0, 1]) < 5 / np.sqrt(B)


def test_merton_multiple_jump_branch(ctx):
    """lam dt = 2.5: counts >= 2 dominate (the second Philox block of jump_size_rare): P(J = 0), the count-mixture KS and the
    first two moments of the compound Poisson law."""
    N, d, B = 4, 1, 500000
    lam = 10.0
    p, dW, J = _merton_increments(ctx, lam, N, d, B, seed=12, muJ=-0.1)
    dt, mu, sj = p["T"] / N, -0.1, p["sigmaJ"]
    m = lam * dt
    n = J.size
    p0 = np.exp(-m)
    assert abs((J == 0).mean() - p0) < 5 * np.sqrt(p0 * (1 - p0) / n)
    assert abs(J.mean() - m * mu) < 5 * np.sqrt(m * (sj ** 2 + mu ** 2) / n)
    assert abs(J.var() / (m * (sj ** 2 + mu ** 2)) - 1) < 0.01
    ks = np.arange(1, 30)
    w = stats.poisson.pmf(ks, m) / (1 - p0)
    cdf = lambda x: sum(wk * stats.norm.cdf(x, k * mu, sj * np.sqrt(k)) for k, wk in zip(ks, w))
    nz = J[J != 0.0][:400000]
    D = ks_distance(nz, cdf)
    print("Merton lam dt = 2.5: KS of the non-zero jumps", D)
    assert D < 1.95 / np.sqrt(nz.size), D


@pytest.mark.parametrize("N,kappa", [(30, 0.1), (5, 0.1), (30, 1.0 / 3.0)], ids=["shape=1/3", "shape=2", "shape=1/10"])
def test_vg_gamma_subordinator(ctx, N, kappa):
    """theta = 1, sigJ -> 0 makes J = gamma: KS against SciPy's Gamma(shape dt/kappa, scale kappa), for the boosted branch
    (shape < 1: mainVG.py's 1/3) and the plain Marsaglia-Tsang branch (shape >= 1)."""
    B = 200000
    par = dict(H.VG, N=N, kappa=kappa, theta=1.0, sigmaJ=1e-20)   # (the sigJ sqrt(gamma) eps term must stay below gamma's own tiny values)
    layout = H.pricing_layout("vg", "SumLocalReg", 1)
    s = H.native_pricing(ctx, "vg", par, "SumLocalReg", layout)
    s.simulate(21, 0, B)
    _, pj, *_ = s.get_noise()
    g = s.read_device(pj, N * B).astype(np.float64)
    shape, scale = (par["T"] / N) / kappa, kappa
    assert abs(g.mean() / (shape * scale) - 1) < 5 * np.sqrt(1.0 / (shape * g.size))
    assert abs(g.var() / (shape * scale ** 2) - 1) < 0.03
    D = ks_distance(g[:400000], lambda x: stats.gamma.cdf(x, shape, scale=scale))
    print(f"VG gamma shape {shape:.3f}: mean {g.mean():.5f} (exact {shape * scale:.5f}), KS {D:.2e}")
    assert D < 1.95 / np.sqrt(min(g.size, 400000)), D


def test_vg_increment_distribution(ctx):
    """mainVG.py parameters: moments of J = theta gamma + sigJ sqrt(gamma) eps and a two-sample KS against the same law drawn
    with NumPy."""
    N, B = 30, 200000
    par = dict(H.VG)
    layout = H.pricing_layout("vg", "SumLocalReg", 1)
    s = H.native_pricing(ctx, "vg", par, "SumLocalReg", layout)
    s.simulate(22, 0, B)
    _, pj, *_ = s.get_noise()
    J = s.read_device(pj, N * B).astype(np.float64)
    dt, th, ka, sj = par["T"] / N, par["theta"], par["kappa"], par["sigmaJ"]
    n = J.size
    var = sj ** 2 * dt + th ** 2 * ka * dt
    assert abs(J.mean() - th * dt) < 5 * np.sqrt(var / n)
    assert abs(J.var() / var - 1) < 0.02
    m3 = (2 * th ** 3 * ka ** 2 + 3 * sj ** 2 * th * ka) * dt        # third central moment of a VG increment
    assert abs(((J - J.mean()) ** 3).mean() / m3 - 1) < 0.1
    rng = np.random.default_rng(5)
    g = rng.gamma(dt / ka, ka, size=600000)
    ref = th * g + sj * np.sqrt(g) * rng.standard_normal(g.size)
    D = stats.ks_2samp(J[:600000], ref).statistic
    print("VG increments: two-sample KS", D)
    assert D < 1.95 * np.sqrt(2.0 / 600000), D


@pytest.mark.parametrize("mean", [1e-3, 0.5, 9.9, 10.1, 25.0, 300.0])
def test_mfg_poisson_sampler_against_scipy_pmf(ctx, mean):
    """Constant-intensity jump model (MFGModel.py:52): dN ~ Poisson(jumpFactor dt) on both sides of the inversion / PTRS switch
    (mean 10) and at the sizes the Cox intensity reaches (1e-3 .. several hundred)."""
    B = 400000
    p = H.mfg_params(1, "constant")
    N = len(p["QAver"]) - 1
    dt = p["T"] / N
    p["jumpFactor"] = mean / dt
    layout = H.mfg_layout("SumLocalReg")
    s = H.native_mfg(ctx, p, "SumLocalReg", layout)
    s.simulate(31, 0, B)
    _, _, pn, *_ = s.get_noise()
    dN = s.read_device(pn, N * B)[:2 * B].astype(np.int64)        # two steps are plenty
    assert (dN >= 0).all()
    n = dN.size
    assert abs(dN.mean() - mean) < 5 * np.sqrt(mean / n)
    assert abs(dN.var() / mean - 1) < 0.02 + 5 * np.sqrt(2.0 / n + 1.0 / (mean * n))
    hi = int(dN.max()) + 1
    emp = np.bincount(dN, minlength=hi + 1) / n
    tv = 0.5 * (np.abs(emp - stats.poisson.pmf(np.arange(hi + 1), mean)).sum() + stats.poisson.sf(hi, mean))
    # expected total-variation distance of an exact sampler ~ 0.4 sqrt(support / n)
    bound = 2.5 * np.sqrt(max(1.0, 8 * np.sqrt(mean)) / n) + 2e-4
    # chi-square goodness of fit over the bins with at least 10 expected counts (the tails are pooled)
    ks = np.arange(hi + 1)
    exp = n * stats.poisson.pmf(ks, mean)
    keep = exp >= 10
    obs_k, exp_k = np.bincount(dN, minlength=hi + 1)[keep].astype(np.float64), exp[keep]
    obs_k = np.append(obs_k, n - obs_k.sum()); exp_k = np.append(exp_k, n - exp_k.sum())
    if exp_k[-1] < 10:
        obs_k[-2] += obs_k[-1]; exp_k[-2] += exp_k[-1]; obs_k, exp_k = obs_k[:-1], exp_k[:-1]
    chi2 = ((obs_k - exp_k) ** 2 / exp_k).sum()
    pval = stats.chi2.sf(chi2, len(exp_k) - 1)
    print(f"Poisson mean {mean}: sample mean {dN.mean():.5f}, var {dN.var():.5f}, TV distance {tv:.2e} (bound {bound:.2e}), "
          f"chi2 {chi2:.1f} on {len(exp_k) - 1} dof (p = {pval:.3f})")
    assert tv < bound and pval > 1e-4


def test_mfg_cox_intensity_is_frozen_at_the_step_start(ctx):
    """Stochastic jump model: replay hQ on the host from the drawn dW0 (MFGModel.py:70) and check the counts against the
    intensity AT THE START of each step (MFGModel.py:47-54): conditional mean and variance lam_i dt, P(dN = 0) = E e^{-lam dt}."""
    B = 300000
    p = H.mfg_params(2, "stochastic")
    Q = np.asarray(p["QAver"], dtype=np.float64)
    N = len(Q) - 1
    dt = p["T"] / N
    layout = H.mfg_layout("SumLocalReg")
    s = H.native_mfg(ctx, p, "SumLocalReg", layout)
    s.simulate(32, 0, B)
    p0, p1, pn, *_ = s.get_noise()
    dW0 = s.read_device(p0, N * B).reshape(N, B).astype(np.float64)
    dW = s.read_device(p1, N * B).reshape(N, B).astype(np.float64)
    dN = s.read_device(pn, N * B).reshape(N, B).astype(np.float64)
    assert abs(dW0.var() / dt - 1) < 5 * np.sqrt(2.0 / dW0.size) and abs(dW.var() / dt - 1) < 5 * np.sqrt(2.0 / dW.size)
    assert abs(np.corrcoef(dW0[3], dW[3])[0, 1]) < 5 / np.sqrt(B)
    hQ = np.full(B, Q[0])
    tot_m = tot_v = tot_z = exp_m = exp_z = 0.0
    means = []
    for i in range(N):
        m = p["beta"] * (np.exp(p["alpha"] * hQ) - 1.0) * dt
        m = np.maximum(m, 0.0)
        means.append(m.mean())
        tot_m += dN[i].sum(); exp_m += m.sum()
        tot_v += ((dN[i] - m) ** 2).sum()
        tot_z += (dN[i] == 0).sum(); exp_z += np.exp(-m).sum()
        hQ = hQ + p["coeffOU"] * (Q[i + 1] - hQ) * dt + p["sig0"] * dW0[i]
    n = N * B
    print(f"Cox counts: step-mean intensity range {min(means):.2e} .. {max(means):.2e}, sum dN / sum lam dt = {tot_m / exp_m:.5f}, "
          f"var ratio {tot_v / exp_m:.5f}, zeros {tot_z / n:.5f} vs {exp_z / n:.5f}")
    assert abs(tot_m / exp_m - 1) < 5 * np.sqrt(1.0 / exp_m)
    assert abs(tot_v / exp_m - 1) < 0.02
    assert abs(tot_z - exp_z) / n < 5 * np.sqrt(0.25 / n)
    # frozen at the START of the step: the innovation dN_i - lam_i dt is uncorrelated with the intensity change over the step;
    # counts drawn at the END-of-step intensity would give S = sum (m_next - m)^2
    hQ = np.full(B, Q[0]); S = var0 = alt = 0.0
    for i in range(N):
        m = (p["beta"] * (np.exp(p["alpha"] * hQ) - 1.0) * dt).clip(0)
        hQ = hQ + p["coeffOU"] * (Q[i + 1] - hQ) * dt + p["sig0"] * dW0[i]
        dm = (p["beta"] * (np.exp(p["alpha"] * hQ) - 1.0) * dt).clip(0) - m
        S += ((dN[i] - m) * dm).sum(); var0 += (m * dm ** 2).sum(); alt += (dm ** 2).sum()
    print(f"frozen-intensity statistic {S:.3f} (sd {np.sqrt(var0):.3f}; end-of-step intensity would give {alt:.3f})")
    assert abs(S) < 5 * np.sqrt(var0) and alt > 10 * np.sqrt(var0)

"""Seeded increment streams shared by tests/golden/make_golden.py (which hands them to the reference) and the tests (which hand
them to the oracle and to the CUDA path).  No imports beyond NumPy: the generator must not depend on the oracle."""
def reg_trajectory_noise(seed, nsteps, N, B, dt, lam, muJ, sigmaJ):
    """Increments of the long Reg-scheme training trajectory (traj/merton_SumLocalReg_300steps.npz): drawn from NumPy's
    frozen MT19937 RandomState stream, so the fixture carries the seed instead of the arrays.  Returns float32 (dW, J), each
    [nsteps, N, B] - the same arrays tests/golden/make_golden.py handed to the reference at its own draw sites."""
    import numpy as np
    rs = np.random.RandomState(int(seed))
    dW = np.empty((nsteps, N, B), dtype=np.float32)
    J = np.empty((nsteps, N, B), dtype=np.float32)
    for k in range(nsteps):
        dW[k] = (np.sqrt(dt) * rs.standard_normal((N, B))).astype(np.float32)
        dN = rs.poisson(lam * dt, (N, B)).astype(np.float64)
        J[k] = (muJ * dN + sigmaJ * np.sqrt(dN) * rs.standard_normal((N, B))).astype(np.float32)
    return dW, J


def vg_trajectory_noise(seed, nsteps, N, B, dt, theta, kappa, sigmaJ):
    """Variance-gamma increments of the long VG Reg trajectory (traj/vg_SumLocalReg_300steps.npz): gamma(shape dt/kappa, scale kappa)
    subordinator steps and the Brownian part on them (pricingModels.py:188-191), float32 [nsteps, N, B], RandomState stream."""
    import numpy as np
    rs = np.random.RandomState(int(seed))
    J = np.empty((nsteps, N, B), dtype=np.float32)
    for k in range(nsteps):
        g = rs.gamma(dt / kappa, kappa, (N, B))
        J[k] = (theta * g + sigmaJ * np.sqrt(g) * rs.standard_normal((N, B))).astype(np.float32)
    return J


def jump_trajectory_noise(kind, seed, nsteps, N, B, M, dt, par):
    """Increments of the long jump-scheme trajectories (traj/*_Global_defaults_*steps.npz; the reference's default shapes: B paths,
    M = 5000 compensator samples redrawn at every time step): float32 (dW or None [nsteps, N, B], J [nsteps, N, B],
    JMC [nsteps, N, M]) from the frozen RandomState stream, drawn step by step in the order dW, J, JMC."""
    import numpy as np
    rs = np.random.RandomState(int(seed))
    merton = kind == "merton"
    dW = np.empty((nsteps, N, B), dtype=np.float32) if merton else None
    J = np.empty((nsteps, N, B), dtype=np.float32)
    JMC = np.empty((nsteps, N, M), dtype=np.float32)

    def jumps(shape):
        if merton:
            dN = rs.poisson(par["lam"] * dt, shape).astype(np.float64)
            return (par["muJ"] * dN + par["sigmaJ"] * np.sqrt(dN) * rs.standard_normal(shape)).astype(np.float32)
        g = rs.gamma(dt / par["kappa"], par["kappa"], shape)
        return (par["theta"] * g + par["sigmaJ"] * np.sqrt(g) * rs.standard_normal(shape)).astype(np.float32)
    for k in range(nsteps):
        if merton:
            dW[k] = (np.sqrt(dt) * rs.standard_normal((N, B))).astype(np.float32)
        J[k] = jumps((N, B))
        JMC[k] = jumps((N, M))
    return dW, J, JMC

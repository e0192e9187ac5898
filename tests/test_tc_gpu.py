"""GPU: the tcgen05 / TMEM plumbing (csrc/tc.cuh) - UMMA descriptors for the K-major and MN-major canonical layouts,
3xTF32 split, TMEM load - against float64 matmuls."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_selftest(ctx):
    import os
    from deepfbsdejsolvers_b200 import _lib as L
    st = C.CDLL(os.path.join(os.path.dirname(L.LIB_PATH), "libfbsdej_selftest.so"))   # test-only kernels live outside the product library
    st.fbsdej_selftest_tc.restype = C.c_int
    st.fbsdej_selftest_tc.argtypes = [C.c_void_p] * 7
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((128, 24)).astype(np.float32), rng.standard_normal((24, 32)).astype(np.float32)
    P, Q = rng.standard_normal((128, 24)).astype(np.float32), rng.standard_normal((128, 24)).astype(np.float32)
    d = [ctx.to_device(x) for x in (A, B, P, Q)]
    o0, o1 = ctx.zeros(4, 128, 32), ctx.zeros(4, 128, 32)
    assert st.fbsdej_selftest_tc(C.c_void_p(ctx.stream.cuda_stream), *[C.c_void_p(t.data_ptr()) for t in d], C.c_void_p(o0.data_ptr()),
                                 C.c_void_p(o1.data_ptr())) == 0
    r0, r1 = ctx.to_host(o0).numpy(), ctx.to_host(o1).numpy()
    ref0 = A.astype(np.float64) @ B.astype(np.float64)
    ref1 = P.astype(np.float64).T @ Q.astype(np.float64)
    e0 = np.abs(r0[0] - ref0).max() / np.abs(ref0).max()
    assert e0 < 5e-6, f"tf32x3 K-major GEMM rel error {e0:.2e}"
    # (tf32 MN-major operands only exist in the SWIZZLE_128B_BASE32B layout: r1[0] is not checked)
    e2 = np.abs(r0[1] - ref0).max() / np.abs(ref0).max()
    e3 = np.abs(r1[1][:24, :24] - ref1).max() / np.abs(ref1).max()
    print("tf32x3 K-major", e0, "bf16x3 K-major", e2, "bf16x3 MN-major", e3)
    assert e2 < 1e-4, f"bf16x3 K-major GEMM rel error {e2:.2e}"
    assert e3 < 1e-4, f"bf16x3 MN-major GEMM rel error {e3:.2e}"
    # A operand in tensor memory (tcgen05.st by the row's own thread), B in shared memory
    e4 = np.abs(r0[2] - ref0).max() / np.abs(ref0).max()
    e5 = np.abs(r0[3] - ref0).max() / np.abs(ref0).max()
    print("A-in-TMEM tf32x3", e4, "bf16x3", e5)
    # M = 64 accumulator: recover the row -> lane map from the dump of all 128 lanes
    m64 = r1[2]
    lanes = []
    for m in range(24):
        d = np.abs(m64[:, :24] - ref1[m][None, :]).max(axis=1) / np.abs(ref1).max()
        lanes.append(int(np.argmin(d)) if d.min() < 1e-4 else -1)
    print("M=64 row->lane", lanes)
    assert e4 < 5e-6, f"A-in-TMEM tf32x3 rel error {e4:.2e}"
    assert e5 < 1e-4, f"A-in-TMEM bf16x3 rel error {e5:.2e}"


import helpers as H
from oracle import MertonOracle, VGOracle


def _check(s, B, l64, g64, g32, aux64, d):
    out, tx, ty, _ = s.loss(B, traj=True)
    assert abs(out[0] - l64) <= 1e-5 * abs(l64), (out[0], l64)
    X = aux64["X"].transpose(0, 2, 1)
    assert np.abs(tx - X).max() <= 2e-6 + 1e-5 * np.abs(X).max()
    Y = aux64["Y"]
    assert np.abs(ty[:Y.shape[0]] - Y).max() <= 4e-6 + 1e-5 * np.abs(Y).max()
    g = s.grad(B)
    assert abs(g[0] - l64) <= 1e-5 * abs(l64)
    scale = np.abs(g64).max()
    e_gpu, e_32 = np.abs(g[4:] - g64).max() / scale, np.abs(g32 - g64).max() / scale
    print("loss rel", abs(out[0] - l64) / abs(l64), "grad rel-to-max", e_gpu, "(fp32 oracle", e_32, ")")
    assert e_gpu <= 5e-5, f"gradient error {e_gpu:.3e}"


@pytest.mark.parametrize("scheme", ["SumLocalReg", "MultiStepReg"])
@pytest.mark.parametrize("d,B", [(1, 300), (10, 1000)])
def test_tensor_core_path_matches_oracle(ctx, scheme, d, B):
    """The tcgen05 kernels (3xTF32 forward, bf16x3 adjoint) against the float64 oracle on injected noise."""
    p = dict(H.MERTON, N=12)
    om = MertonOracle(aLin=H.ALIN, limit=30 if d == 1 else 100, d=d, **p)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 21)
    noise = H.merton_noise(om, B, 0, seed=22, with_jmc=False)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    # d = 10 through the per-step Hermite table of the closed form (the default of the model class), d = 1 through the series
    s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, limit=30 if d == 1 else 100, tensor_cores=True, price_table=d > 1)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), None)
    _check(s, B, l64, g64, g32, aux64, d)


def test_tensor_core_path_vg(ctx):
    B, scheme = 500, "SumLocalReg"
    om = VGOracle(aLin=H.ALIN, **H.VG)
    layout = H.pricing_layout("vg", scheme, 1)
    theta = H.random_theta(layout, 23)
    noise = H.vg_noise(om, B, 0, seed=24, with_jmc=False)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "vg", H.VG, scheme, layout, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, None, H.to_planes(noise["J"]), None)
    _check(s, B, l64, g64, g32, aux64, 1)


@pytest.mark.parametrize("act,B,Hn", [("relu", 37, 21), ("tanh", 129, 16), ("relu", 260, 22)])
def test_tensor_core_path_shapes_and_relu(ctx, act, B, Hn):
    """Edge shapes of the tcgen05 kernels: fewer paths than one 128-row tile, one path into the second tile, the widest and a
    narrow hidden layer, ReLU (whose constant-1 unit and masks take the other branch of the kernels)."""
    d, scheme = 10, "SumLocalReg"
    p = dict(H.MERTON, N=6)
    om = MertonOracle(aLin=H.ALIN, limit=100, d=d, **p)
    layout = H.pricing_layout("merton", scheme, d, H=Hn, act=act)
    theta = H.random_theta(layout, 31)
    noise = H.merton_noise(om, B, 0, seed=32, with_jmc=False)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, limit=100, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), None)
    _check(s, B, l64, g64, g32, aux64, d)


from oracle import MFGOracle


@pytest.mark.parametrize("scheme", ["Global", "MultiStep", "SumLocal", "SumLocalReg", "MultiStepReg"])
@pytest.mark.parametrize("jumpModel,B", [("stochastic", 130), ("constant", 64)])
def test_mfg_tensor_core_path_matches_oracle(ctx, scheme, jumpModel, B):
    """The tcgen05 MFG kernels (two networks side by side, nout up to 4) against the float64 oracle on injected increments."""
    from oracle.mfg import sample_mfg_noise
    p = H.mfg_params(1, jumpModel)
    om = MFGOracle(**p)
    layout = H.mfg_layout(scheme)
    theta = H.random_theta(layout, 4)
    noise = sample_mfg_noise(om, B, torch.Generator().manual_seed(8))
    (lh32, li32), g32, _ = H.oracle_mfg(om, scheme, layout, theta, noise, B)
    (lh64, li64), g64, aux64 = H.oracle_mfg(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_mfg(ctx, p, scheme, layout, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, noise["dW0"].numpy(), noise["dW"].numpy(), noise["dN"].numpy())
    out, tx, ty, _ = s.loss(B, traj=True)
    assert abs(out[1] - lh64) <= 1e-5 * abs(lh64) and abs(out[2] - li64) <= 1e-5 * abs(li64), (out, lh64, li64)
    assert np.abs(tx[:, 0, :] - aux64["hS"]).max() <= 2e-5 and np.abs(tx[:, 1, :] - aux64["S"]).max() <= 2e-5
    g = s.grad(B)
    scale = np.abs(g64).max()
    e_gpu, e_32 = np.abs(g[4:] - g64).max() / scale, np.abs(g32 - g64).max() / scale
    print(scheme, jumpModel, "loss rel", abs(out[0] - lh64 - li64) / abs(lh64 + li64), "grad rel-to-max", e_gpu, "(fp32 oracle", e_32, ")")
    assert e_gpu <= 5e-5, f"gradient error {e_gpu:.3e}"


# Gradient bound of the jump schemes on tcgen05: the parameter gradient of the jump network is Ybar (dG(own jump) - mean_m dG(sample m)),
# a difference of two nearly equal sums (VG: |h_own - mean h| ~ 0.05 |h|), so the 2^-17 relative rounding of the bf16x3 products
# is amplified ~20x: measured up to 2.3e-4 of the largest component (VG Global, 24 paths x 300 samples), 1e-6 on the fp32 FFMA
# kernels (mma_mode = 0).  Losses and trajectories keep 1e-5.
JUMP_TC_GRAD_TOL = 3e-4


def _check_jump(s, B, l64, g64, g32, aux64, has_z):
    out, tx, ty, tz = s.loss(B, traj=True)
    assert abs(out[0] - l64) <= 1e-5 * abs(l64), (out[0], l64)
    X = aux64["X"][:, :, 0]
    assert np.abs(tx[:, 0, :] - X).max() <= 2e-6 + 1e-5 * np.abs(X).max()
    Y = aux64["Y"]
    assert np.abs(ty[:Y.shape[0]] - Y).max() <= 4e-6 + 1e-5 * np.abs(Y).max()
    g = s.grad(B)
    assert abs(g[0] - l64) <= 1e-5 * abs(l64)
    scale = np.abs(g64).max()
    e_gpu, e_32 = np.abs(g[4:] - g64).max() / scale, np.abs(g32 - g64).max() / scale
    print("loss rel", abs(out[0] - l64) / abs(l64), "grad rel-to-max", e_gpu, "(fp32 oracle", e_32, ")")
    assert e_gpu <= JUMP_TC_GRAD_TOL, f"gradient error {e_gpu:.3e}"


@pytest.mark.parametrize("scheme", ["Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2"])
@pytest.mark.parametrize("B,M", [(10, 700), (37, 160), (1500, 40)])
def test_jump_network_on_tensor_cores_merton(ctx, scheme, B, M):
    """Jump schemes with the jump evaluations (own jump + Monte-Carlo compensator rows) on tcgen05: a cluster of CTAs
    per path (B = 10), one CTA per path, and several paths per 128-row tile (B = 1500)."""
    om = MertonOracle(aLin=H.ALIN, limit=30, d=1, **H.MERTON)
    layout = H.pricing_layout("merton", scheme, 1)
    theta = H.random_theta(layout, 41)
    noise = H.merton_noise(om, B, M, seed=42, with_jmc=True)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", H.MERTON, scheme, layout, d=1, M=M, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]))
    _check_jump(s, B, l64, g64, g32, aux64, True)


@pytest.mark.parametrize("scheme", ["Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2"])
def test_jump_network_on_tensor_cores_vg(ctx, scheme):
    B, M = 24, 300
    om = VGOracle(aLin=H.ALIN, **H.VG)
    layout = H.pricing_layout("vg", scheme, 1)
    theta = H.random_theta(layout, 43)
    noise = H.vg_noise(om, B, M, seed=44, with_jmc=True)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "vg", H.VG, scheme, layout, M=M, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, None, H.to_planes(noise["J"]), H.to_planes(noise["JMC"]))
    _check_jump(s, B, l64, g64, g32, aux64, False)


@pytest.mark.parametrize("scheme", ["Global", "MultiStep1", "MultiStep2", "SumLocal1", "SumLocal2"])
@pytest.mark.parametrize("B,M", [(64, 400), (1500, 96)])
def test_jump_network_on_tensor_cores_merton_d10(ctx, scheme, B, M):
    """d = 10: the jump rows have 22 input features (three 8-feature chunks of the operand tile); one path per CTA with a
    cluster (B = 64) and several paths per tile (B = 1500)."""
    d = 10
    p = dict(H.MERTON, N=12)
    om = MertonOracle(aLin=H.ALIN, limit=100, d=d, **p)
    layout = H.pricing_layout("merton", scheme, d)
    theta = H.random_theta(layout, 51)
    noise = H.merton_noise(om, B, M, seed=52, with_jmc=True)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", p, scheme, layout, d=d, M=M, limit=100, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]))
    out, tx, ty, tz = s.loss(B, traj=True)
    assert abs(out[0] - l64) <= 1e-5 * abs(l64), (out[0], l64)
    X = aux64["X"].transpose(0, 2, 1)
    assert np.abs(tx - X).max() <= 2e-6 + 1e-5 * np.abs(X).max()
    g = s.grad(B)
    scale = np.abs(g64).max()
    e_gpu, e_32 = np.abs(g[4:] - g64).max() / scale, np.abs(g32 - g64).max() / scale
    print("loss rel", abs(out[0] - l64) / abs(l64), "grad rel-to-max", e_gpu, "(fp32 oracle", e_32, ")")
    assert e_gpu <= JUMP_TC_GRAD_TOL, f"gradient error {e_gpu:.3e}"


def test_jump_network_on_tensor_cores_one_path_per_thread(ctx):
    """Large batch: pick_G gives every thread its own path (G = 1), a tile is 128 paths on the same compensator sample."""
    B, M, scheme = 40000, 24, "Global"
    p = dict(H.MERTON, N=6)
    om = MertonOracle(aLin=H.ALIN, limit=30, d=1, **p)
    layout = H.pricing_layout("merton", scheme, 1)
    theta = H.random_theta(layout, 61)
    noise = H.merton_noise(om, B, M, seed=62, with_jmc=True)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, "merton", p, scheme, layout, d=1, M=M, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]), H.to_planes(noise["J"]), H.to_planes(noise["JMC"]))
    _check_jump(s, B, l64, g64, g32, aux64, True)


@pytest.mark.parametrize("kind", ["reg", "jump", "mfg"])
def test_tensor_core_kernels_are_bitwise_repeatable(ctx, kind):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer_unavailable.txt), so the hand-written synchronisation of
    the tcgen05 kernels (operand-tile reuse behind mbarriers, the warp-specialised ring, the DSMEM cluster sums) is exercised the
    other way round: a race shows up as run-to-run differences, and eight repetitions of the same call must agree bit for bit."""
    from oracle.mfg import sample_mfg_noise
    if kind == "mfg":
        p = H.mfg_params(1, "stochastic")
        layout = H.mfg_layout("MultiStep")
        s = H.native_mfg(ctx, p, "MultiStep", layout, tensor_cores=True)
        s.set_theta(H.random_theta(layout, 4))
        nz = sample_mfg_noise(MFGOracle(**p), 300, torch.Generator().manual_seed(8))
        s.set_noise(300, nz["dW0"].numpy(), nz["dW"].numpy(), nz["dN"].numpy())
        run = lambda: s.grad(300)
    elif kind == "reg":
        p = dict(H.MERTON, N=25)
        layout = H.pricing_layout("merton", "SumLocalReg", 10)
        s = H.native_pricing(ctx, "merton", p, "SumLocalReg", layout, d=10, limit=100, tensor_cores=True, price_table=True)
        s.set_theta(H.random_theta(layout, 5))
        run = lambda: ctx.to_host(s.grad_step(77, 70001, 70001, 0)).numpy().copy()     # mixed 128 / 96-row tiles, in-kernel increments
    else:
        p = dict(H.MERTON, N=10)
        layout = H.pricing_layout("merton", "Global", 1)
        s = H.native_pricing(ctx, "merton", p, "Global", layout, d=1, M=500, tensor_cores=True)
        s.set_theta(H.random_theta(layout, 6))
        s.simulate(5, 0, 12)                                                           # clusters of CTAs per path
        run = lambda: s.grad(12)
    first = run()
    assert np.isfinite(first).all()
    for _ in range(7):
        assert np.array_equal(run(), first)


@pytest.mark.parametrize("kind,scheme,d,M", [("merton", "Global", 1, 24), ("merton", "MultiStep2", 1, 300), ("merton", "SumLocal2", 10, 40),
                                             ("merton", "Global", 10, 150), ("vg", "Global", 1, 40), ("vg", "SumLocal2", 1, 200)])
def test_jump_network_separable_first_layer(ctx, kind, scheme, d, M):
    """One path per thread (large batch, two-network schemes): the jump rows of a tile share their compensator sample, and the first
    layer splits into a path part and a sample part (jump_tc.cuh: preact / eval_sep and their adjoint) - against the float64 oracle,
    and against the same kernels with the split turned off."""
    B = 38000
    if kind == "merton":
        p = dict(H.MERTON, N=3)
        om = MertonOracle(aLin=H.ALIN, limit=30 if d == 1 else 100, d=d, **p)
        noise = H.merton_noise(om, B, M, seed=92, with_jmc=True)
    else:
        p = dict(H.VG, N=3)
        om = VGOracle(aLin=H.ALIN, **p)
        noise = H.vg_noise(om, B, M, seed=92, with_jmc=True)
    layout = H.pricing_layout(kind, scheme, d)
    theta = H.random_theta(layout, 91)
    l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
    l64, g64, aux64 = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
    s = H.native_pricing(ctx, kind, p, scheme, layout, d=d, M=M, limit=30 if d == 1 else 100, tensor_cores=True)
    s.set_theta(theta)
    s.set_noise(B, H.to_planes(noise["dW"]) if "dW" in noise else None, H.to_planes(noise["J"]), H.to_planes(noise["JMC"]))
    out, tx, ty, tz = s.loss(B, traj=True)
    assert abs(out[0] - l64) <= 1e-5 * abs(l64), (out[0], l64)
    X = aux64["X"].transpose(0, 2, 1)
    assert np.abs(tx - X).max() <= 2e-6 + 1e-5 * np.abs(X).max()
    g = s.grad(B)
    scale = np.abs(g64).max()
    e_gpu = np.abs(g[4:] - g64).max() / scale
    print(kind, scheme, d, "loss rel", abs(out[0] - l64) / abs(l64), "grad rel-to-max", e_gpu)
    assert e_gpu <= JUMP_TC_GRAD_TOL, f"gradient error {e_gpu:.3e}"

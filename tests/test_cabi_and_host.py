"""CPU: the C-ABI library loads and exports every symbol include/fbsdej.h declares (no compute calls), plus the host
logic of the drop-in layer (parameter layout, sharding, coupling recognition, initialisers)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fbsdej.h")).read()
    return re.findall(r"FBSDEJ_API\s+[\w\s\*]+?\b(fbsdej_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    import deepfbsdejsolvers_b200 as pkg
    syms = header_symbols()
    assert len(syms) >= 30
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/fbsdej.h but not exported"
    assert set(pkg.SYMBOLS) == set(syms), set(pkg.SYMBOLS) ^ set(syms)
    assert lib.fbsdej_version() == 100


def test_error_path_without_gpu():
    import torch
    import deepfbsdejsolvers_b200 as pkg
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.FbsdejError):
        pkg.Context()


def test_oracle_is_not_imported_by_the_product():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deepfbsdejsolvers_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_reference_parameter_counts():
    """2->21->21->{1,2} = 547/569, 3->21->21->1 = 568; MFG 4->20->20->{1,2,3}, 6->22->22->{1,3,4}."""
    from deepfbsdejsolvers_b200 import NetSpec
    assert NetSpec(2, 21, 1).nparams == 547 and NetSpec(2, 21, 2).nparams == 569 and NetSpec(3, 21, 1).nparams == 568
    assert [NetSpec(4, 20, k).nparams for k in (1, 2, 3)] == [541, 562, 583]
    assert [NetSpec(6, 22, k).nparams for k in (1, 3, 4)] == [683, 729, 752]


def test_net_lazy_build_and_layout():
    from deepfbsdejsolvers_b200 import set_seed
    from deepfbsdejsolvers_b200.coupledPricing import Net
    from deepfbsdejsolvers_b200.coupledMFG import kerasModels, Net_hat, Net as NetM
    set_seed(3)
    n = Net(1, 1, 21 * np.ones((2,), dtype=np.int32), "tanh")
    assert n.params is None and hasattr(n, "Y0")
    n.build(2)
    assert n.params.size == 547 and n.params.dtype == np.float32
    (W1, b1), (W2, b2), (W3, b3) = n.layer_arrays()
    assert W1.shape == (2, 21) and W2.shape == (21, 21) and W3.shape == (21, 1)
    assert not b1.any() and not b2.any() and not b3.any()
    std = np.sqrt(2.0 / 42) / 0.87962566
    assert np.abs(W2).max() <= 2 * std + 1e-6 and 0.5 * std < W2.std() < 1.2 * std       # truncated Glorot normal
    with pytest.raises(ValueError):
        n.build(3)
    assert not hasattr(Net(0, 2, [21, 21], "relu"), "Y0")
    km = kerasModels(Net_hat, NetM, "Global", 2, 3, [20, 20], [22, 22], "tanh", "tanh")
    assert hasattr(km.model_hat, "Y0_hat") and hasattr(km.model, "Y0")
    assert not hasattr(kerasModels(Net_hat, NetM, "SumLocal", 3, 4, [20, 20], [22, 22], "tanh", "tanh").model, "Y0")
    n3 = Net(0, 1, [21, 21, 21], "tanh")             # nbLayer = 3 (mainMerton.py:13): three equal hidden layers
    n3.build(2)
    assert n3.spec().L == 3 and n3.params.size == 2 * 21 + 21 + 2 * (21 * 21 + 21) + 21 + 1 and len(n3.layer_arrays()) == 4
    assert Net(0, 1, [16], "relu").L == 1
    with pytest.raises(ValueError):
        Net(0, 1, [21, 21, 21, 21], "tanh")
    with pytest.raises(ValueError):
        Net(0, 1, [21, 16], "tanh")


def test_coupling_recognition():
    import torch
    from deepfbsdejsolvers_b200.coupledPricing.pricingModels import coupling_slope, AbsCoupling
    assert coupling_slope(AbsCoupling(0.1)) == 0.1
    assert abs(coupling_slope(lambda x: 0.25 * torch.abs(x)) - 0.25) < 1e-12
    assert abs(coupling_slope(lambda x: 0.1 * abs(x)) - 0.1) < 1e-12
    with pytest.raises(ValueError):
        coupling_slope(lambda x: x * x)


def test_shard_partition():
    from deepfbsdejsolvers_b200.solver_base import shard
    for B in (10, 128, 2 ** 20, 1000003):
        for world in (1, 2, 4, 8):
            parts = [shard(B, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == B
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_mfg_model_host_api_matches_oracle():
    """The drop-in ModelCoupledFBSDE per-step API (host tensors) against the oracle's restatement."""
    import torch
    import helpers as H
    from oracle import MFGOracle
    from deepfbsdejsolvers_b200.coupledMFG import ModelCoupledFBSDE
    p = H.mfg_params(1)
    mm, om = ModelCoupledFBSDE(**p), MFGOracle(**p)
    assert mm.N == om.N == 47
    B = 5
    mm.init(B)
    st = om.init(B)
    g = torch.Generator().manual_seed(0)
    for i in range(6):
        dW0, dW = 0.1 * torch.randn(B, generator=g), 0.1 * torch.randn(B, generator=g)
        dN = (torch.rand(B, generator=g) < 0.3).float()
        hY, Y = torch.randn(B, generator=g), torch.randn(B, generator=g)
        np.testing.assert_allclose(mm.calpha_hat(hY).numpy(), om.calpha_hat(st, hY).numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(mm.calpha(hY, Y).numpy(), om.calpha(st, hY, Y).numpy(), rtol=1e-5, atol=1e-6)
        mm.oneStepFrom(dW0, dW, dN, hY, Y)
        st = om.one_step(st, dW0, dW, dN, hY, Y)
        for a, b in ((mm.hQ, st["hQ"]), (mm.Q, st["Q"]), (mm.R, st["R"]), (mm.hS, st["hS"]), (mm.S, st["S"])):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-6)
    assert abs(mm.meanhQ - om.mean_hq(6)) < 1e-12

"""GPU: the tcgen05 / TMEM plumbing (csrc/tc.cuh) - UMMA descriptors for the K-major and MN-major canonical layouts,
3xTF32 split, TMEM load - against float64 matmuls."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_selftest(ctx):
    from deepfbsdejsolvers_b200 import _lib as L
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((128, 24)).astype(np.float32), rng.standard_normal((24, 32)).astype(np.float32)
    P, Q = rng.standard_normal((128, 24)).astype(np.float32), rng.standard_normal((128, 24)).astype(np.float32)
    d = [ctx.to_device(x) for x in (A, B, P, Q)]
    o0, o1 = ctx.zeros(2, 128, 32), ctx.zeros(2, 128, 32)
    L.check(L.lib.fbsdej_selftest_tc(ctx.handle, *[C.c_void_p(t.data_ptr()) for t in d], C.c_void_p(o0.data_ptr()),
                                     C.c_void_p(o1.data_ptr())))
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

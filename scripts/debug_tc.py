import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from deepfbsdejsolvers_b200 import Context, _lib as L
ctx = Context.default()
rng = np.random.default_rng(0)
A, B = rng.standard_normal((128, 24)).astype(np.float32), rng.standard_normal((24, 32)).astype(np.float32)
P, Q = rng.standard_normal((128, 24)).astype(np.float32), rng.standard_normal((128, 24)).astype(np.float32)
d = [ctx.to_device(x) for x in (A, B, P, Q)]
o0, o1 = ctx.zeros(128, 32), ctx.zeros(128, 32)
L.check(L.lib.fbsdej_selftest_tc(ctx.handle, *[C.c_void_p(t.data_ptr()) for t in d], C.c_void_p(o0.data_ptr()), C.c_void_p(o1.data_ptr())))
np.savez("gpurun_out/tc_debug.npz", A=A, B=B, P=P, Q=Q, r0=ctx.to_host(o0).numpy(), r1=ctx.to_host(o1).numpy())
print("saved")

import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers as H
from deepfbsdejsolvers_b200 import Context
ctx = Context.default(0)
P = H.mfg_params(2)
for tc in (False, True):
    for scheme in ("Global", "SumLocalReg"):
        layout = H.mfg_layout(scheme)
        s = H.native_mfg(ctx, P, scheme, layout, tensor_cores=tc)
        s.set_theta(H.random_theta(layout, 1))
        print("tc", tc, scheme, {k: round(v, 4) for k, v in s.profile(0, 128, reps=20).items()})

import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
import helpers as H
from oracle import VGOracle
from deepfbsdejsolvers_b200 import Context
ctx = Context.default()
for scheme in ["Global", "MultiStep2"]:
    for B, M in [(24, 128), (96, 128), (24, 16)]:
        om = VGOracle(aLin=H.ALIN, **H.VG)
        layout = H.pricing_layout("vg", scheme, 1)
        theta = H.random_theta(layout, 3)
        noise = H.vg_noise(om, B, M, seed=7)
        l32, g32, _ = H.oracle_pricing(om, scheme, layout, theta, noise, B)
        l64, g64, aux = H.oracle_pricing(om, scheme, layout, theta, noise, B, dtype=torch.float64)
        s = H.native_pricing(ctx, "vg", H.VG, scheme, layout, M=M)
        s.set_theta(theta)
        s.set_noise(B, None, H.to_planes(noise["J"]), H.to_planes(noise["JMC"]))
        g = s.grad(B)[4:]
        sc = np.abs(g64).max()
        e = np.abs(g - g64) / sc
        e32 = np.abs(g32 - g64) / sc
        idx = np.argsort(-e)[:5]
        print(scheme, B, M, "max err", e.max(), "fp32 oracle", e32.max(), "argmax|g|", np.argmax(np.abs(g64)), "P", g.size,
              "offsets", layout.offsets, layout.y0_offset)
        for i in idx:
            print("   idx", i, "gpu", g[i], "f64", g64[i], "f32", g32[i])
        print("   J range", float(noise["J"].min()), float(noise["J"].max()), "X range", aux["X"].min(), aux["X"].max())

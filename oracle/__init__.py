"""CPU oracle for the deep-FBSDE-with-jumps training iteration.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a torch-CPU (autograd) restatement of the reference algorithm
(ZakariaBensaid/DeepFBSDEJSolvers, files cited per function) with *injectable
noise*, used only as the checker by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.  Nothing under
`deepfbsdejsolvers_b200/` imports it.

Parity pin: the reference ships no tests or golden vectors and TensorFlow cannot
be imported in this image.  The oracle is pinned instead against
  (a) the reference's own source files executed unmodified through a
      torch-backed stand-in for the `tensorflow` API (tests/golden/tfshim,
      generator tests/golden/make_golden.py, fixtures tests/golden/*.npz), and
  (b) the closed-form known answers embedded in the reference
      (Merton series 0.2714569268, VG FFT price 0.1331402194).
TensorFlow-internal semantics (Dense, Glorot initialisers, Keras Adam, abs/max
sub-gradients) are restated from their published definitions; they are the part
of the pin that is *not* executed reference code.
"""
from .nets import MLPSpec, ParamLayout, glorot_normal, glorot_uniform, mlp_forward  # noqa: F401
from .adam import KerasAdam  # noqa: F401
from .pricing import MertonOracle, VGOracle, pricing_loss, PRICING_SCHEMES  # noqa: F401
from .mfg import MFGOracle, mfg_loss, MFG_SCHEMES  # noqa: F401

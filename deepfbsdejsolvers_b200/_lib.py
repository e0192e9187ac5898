"""ctypes binding of libfbsdej.so (C-ABI declared in include/fbsdej.h).

The product path has NO CPU fallback: importing this module without the built CUDA library raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FBSDEJ_LIB") or os.path.join(_HERE, "libfbsdej.so")   # FBSDEJ_LIB: another build of the same library

MODEL_MERTON, MODEL_VG, MODEL_MFG = 0, 1, 2
GLOBAL, MULTISTEP1, MULTISTEP2, SUMLOCAL1, SUMLOCAL2, SUMLOCALREG, MULTISTEPREG = range(7)
ACT = {"tanh": 0, "relu": 1}
OUT_HEADER = 4


class FbsdejError(RuntimeError):
    pass


class MertonParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("T", "r", "muJ", "sigJ", "sig", "lam", "K", "x0", "aLin")] + \
               [(n, C.c_int) for n in ("N", "limit", "d")]


class VGParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("T", "r", "theta", "kappa", "sigJ", "K", "x0", "aLin")] + [("N", C.c_int)]


class MFGParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("T", "R0", "jumpFactor", "alpha", "beta", "coeffOU", "A", "K", "pi", "p0", "p1",
                                          "f0", "f1", "theta", "C", "S0", "h1", "h2", "sig0", "sig", "alphaTarget",
                                          "coeffEqui")] + \
               [("stochastic_jumps", C.c_int), ("nQ", C.c_int), ("QAver", C.POINTER(C.c_double))]


class NetDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nin", "nout", "H", "L", "act")]


class SolverDesc(C.Structure):
    _fields_ = [("model", C.c_int), ("scheme", C.c_int), ("n_nets", C.c_int), ("nets", NetDesc * 2), ("n_y0", C.c_int),
                ("M", C.c_int), ("stale_time", C.c_int), ("w_hat", C.c_float), ("w_ind", C.c_float), ("price_table", C.c_int), ("mma_mode", C.c_int)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise FbsdejError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C deepfbsdejsolvers_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
    sig = {
        "fbsdej_last_error": (C.c_char_p, []),
        "fbsdej_version": (i32, []),
        "fbsdej_ctx_create": (i32, [i32, vp, C.POINTER(vp)]),
        "fbsdej_ctx_destroy": (i32, [vp]),
        "fbsdej_ctx_sync": (i32, [vp]),
        "fbsdej_malloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "fbsdej_free": (i32, [vp, vp]),
        "fbsdej_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t]),
        "fbsdej_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "fbsdej_solver_create": (i32, [vp, C.POINTER(SolverDesc), C.POINTER(MertonParams), C.POINTER(VGParams),
                                       C.POINTER(MFGParams), C.POINTER(vp)]),
        "fbsdej_solver_destroy": (i32, [vp]),
        "fbsdej_solver_nparams": (i32, [vp]),
        "fbsdej_solver_set_weights": (i32, [vp, f32, f32]),
        "fbsdej_solver_set_vg_table_host": (i32, [vp, C.POINTER(C.c_double), i32, C.c_double, C.c_double]),
        "fbsdej_solver_simulate": (i32, [vp, u64, u32, u32, i32]),
        "fbsdej_solver_set_noise": (i32, [vp, i32, vp, vp, vp]),
        "fbsdej_solver_set_noise_sparse_jumps": (i32, [vp, i32, vp, vp, vp, i32, vp]),
        "fbsdej_solver_get_noise": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "fbsdej_solver_loss": (i32, [vp, vp, i32, i32, vp, vp, vp, vp]),
        "fbsdej_solver_mfg_states": (i32, [vp, i32, vp]),
        "fbsdej_solver_grad": (i32, [vp, vp, i32, i32, vp]),
        "fbsdej_adam_step": (i32, [vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, vp]),
        "fbsdej_solver_grad_step": (i32, [vp, vp, u64, vp, u32, i32, i32, vp]),
        "fbsdej_bump_u32": (i32, [vp, vp]),
        "fbsdej_solver_train_steps": (i32, [vp, vp, vp, vp, vp, vp, vp, u64, i32, i32, f32, f32, f32, f32, vp]),
        "fbsdej_solver_dp_init": (i32, [vp, i32, i32, vp]),
        "fbsdej_solver_dp_buffer": (i32, [vp, vp]),
        "fbsdej_solver_dp_connect": (i32, [vp, vp, vp]),
        "fbsdej_solver_dp_check": (i32, [vp]),
        "fbsdej_solver_train_steps_dp": (i32, [vp, vp, vp, vp, vp, vp, vp, u64, i32, i32, u32, i32, f32, f32, f32, f32, vp]),
        "fbsdej_solver_profile": (i32, [vp, vp, u64, i32, i32, C.POINTER(C.c_float)]),
        "fbsdej_solver_net_forward": (i32, [vp, vp, i32, vp, i32, vp]),
        "fbsdej_net_forward": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp]),
        "fbsdej_solver_price": (i32, [vp, i32, vp, i32, vp]),
        "fbsdej_transpose_nbd_to_ndb": (i32, [vp, vp, vp, i32, i32, i32]),
        "fbsdej_transpose_ndb_to_nbd": (i32, [vp, vp, vp, i32, i32, i32]),
        "fbsdej_ctx_launch_count": (C.c_longlong, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sig)


lib, SYMBOLS = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise FbsdejError(f"fbsdej error {rc}: {lib.fbsdej_last_error().decode()}")

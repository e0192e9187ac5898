"""Host-side runtime above the C-ABI: one `Context` per GPU/rank, one `NativeSolver` per (model, scheme, nets).

PyTorch is used only as the carrier of device memory, streams and (for data parallelism) torch.distributed;
every kernel is launched by libfbsdej.so.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from ._lib import FbsdejError, check, lib


@dataclass
class NetSpec:
    nin: int
    H: int
    nout: int
    activation: str = "tanh"
    L: int = 2

    @property
    def nparams(self) -> int:
        return self.nin * self.H + self.H + (self.L - 1) * (self.H * self.H + self.H) + self.H * self.nout + self.nout


class Context:
    """fbsdej_ctx + a dedicated (non-default) CUDA stream; not thread-safe, one per GPU."""

    _default = {}

    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise FbsdejError("no CUDA device visible: deepfbsdejsolvers_b200 has no CPU fallback")
        idx = torch.cuda.current_device() if device is None else int(device)
        self.index = idx
        self.device = torch.device("cuda", idx)
        self.stream = torch.cuda.Stream(self.device)
        h = C.c_void_p()
        check(lib.fbsdej_ctx_create(idx, C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.handle = h

    @classmethod
    def default(cls, device: Optional[int] = None) -> "Context":
        idx = torch.cuda.current_device() if device is None else int(device)
        if idx not in cls._default:
            cls._default[idx] = cls(idx)
        return cls._default[idx]

    def sync(self) -> None:
        check(lib.fbsdej_ctx_sync(self.handle))

    @property
    def launches(self) -> int:
        return int(lib.fbsdej_ctx_launch_count(self.handle))

    def empty(self, *shape, dtype=torch.float32) -> torch.Tensor:
        with torch.cuda.stream(self.stream):
            return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32) -> torch.Tensor:
        with torch.cuda.stream(self.stream):
            return torch.zeros(*shape, dtype=dtype, device=self.device)

    def to_device(self, x, dtype=torch.float32) -> torch.Tensor:
        t = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
        with torch.cuda.stream(self.stream):
            return t.to(device=self.device, dtype=dtype, non_blocking=False).contiguous()

    def to_host(self, t: torch.Tensor) -> torch.Tensor:
        with torch.cuda.stream(self.stream):
            out = t.detach().to("cpu")
        self.stream.synchronize()
        return out

    def net_forward(self, theta_net: torch.Tensor, spec: NetSpec, x: torch.Tensor) -> torch.Tensor:
        """Rows of x [rows, nin] through one network -> [rows, nout] (Net.call, Networks.py:17-23)."""
        rows = x.shape[0]
        y = self.empty(rows, spec.nout)
        check(lib.fbsdej_net_forward(self.handle, _p(theta_net), spec.nin, spec.H, spec.L, spec.nout, L.ACT[spec.activation],
                                     _p(x), rows, _p(y)))
        return y


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


class NativeSolver:
    """fbsdej_solver + its flat parameter / Adam state on the device."""

    def __init__(self, ctx: Context, model_kind: int, scheme: int, nets: Sequence[NetSpec], n_y0: int, M: int = 0,
                 merton: Optional[L.MertonParams] = None, vg: Optional[L.VGParams] = None,
                 mfg: Optional[L.MFGParams] = None, stale_time: bool = True, w_hat: float = 1.0, w_ind: float = 1.0,
                 price_table: bool = False, tensor_cores: bool = False):
        self.ctx, self.model_kind, self.scheme, self.nets, self.n_y0, self.M = ctx, model_kind, scheme, list(nets), n_y0, M
        d = L.SolverDesc()
        d.model, d.scheme, d.n_nets, d.n_y0, d.M = model_kind, scheme, len(nets), n_y0, M
        d.stale_time, d.w_hat, d.w_ind, d.price_table = int(stale_time), w_hat, w_ind, int(price_table)
        d.mma_mode = int(tensor_cores)
        for k, n in enumerate(nets):
            d.nets[k] = L.NetDesc(n.nin, n.nout, n.H, n.L, L.ACT[n.activation])
        h = C.c_void_p()
        check(lib.fbsdej_solver_create(ctx.handle, C.byref(d), C.byref(merton) if merton is not None else None,
                                       C.byref(vg) if vg is not None else None,
                                       C.byref(mfg) if mfg is not None else None, C.byref(h)))
        self.handle = h
        self.P = int(lib.fbsdej_solver_nparams(h))
        self.offsets: List[int] = []
        off = 0
        for n in nets:
            self.offsets.append(off)
            off += n.nparams
        self.y0_offset = off
        assert off + n_y0 == self.P
        self.theta = ctx.zeros(self.P)
        self.m = ctx.zeros(self.P)
        self.v = ctx.zeros(self.P)
        self.t = ctx.zeros(1, dtype=torch.int32)
        self.iteration = ctx.zeros(1, dtype=torch.int32)     # uint32 Philox iteration word
        self.out = ctx.zeros(L.OUT_HEADER + self.P)
        self._keep = []                                      # tensors the C side holds pointers to

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.fbsdej_solver_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------------------------------
    def set_theta(self, theta) -> None:
        t = self.ctx.to_device(theta)
        assert t.numel() == self.P
        with torch.cuda.stream(self.ctx.stream):
            self.theta.copy_(t.reshape(-1))

    def get_theta(self) -> np.ndarray:
        return self.ctx.to_host(self.theta).numpy()

    def reset_optimizer(self) -> None:
        with torch.cuda.stream(self.ctx.stream):
            self.m.zero_(); self.v.zero_(); self.t.zero_()

    # ---- checkpoint (params | m | v | t | iteration; the reference keeps nothing on disk) ---------------
    def state_dict(self) -> dict:
        """Everything a bit-exact resume needs: the Philox draws depend only on (seed, iteration, path id)."""
        h = self.ctx.to_host
        return {"theta": h(self.theta).numpy().copy(), "m": h(self.m).numpy().copy(), "v": h(self.v).numpy().copy(),
                "t": h(self.t).numpy().copy(), "iteration": h(self.iteration).numpy().copy(),
                "layout": np.array([self.model_kind, self.scheme, self.P, self.n_y0, self.M], dtype=np.int64)}

    def load_state_dict(self, sd: dict) -> None:
        lay = np.asarray(sd["layout"]).astype(np.int64)
        mine = np.array([self.model_kind, self.scheme, self.P, self.n_y0, self.M], dtype=np.int64)
        if not np.array_equal(lay, mine):
            raise FbsdejError(f"checkpoint layout {lay.tolist()} does not match this solver {mine.tolist()} "
                              "(model, scheme, parameter count, Y0 count, compensator samples)")
        with torch.cuda.stream(self.ctx.stream):
            self.theta.copy_(torch.from_numpy(np.asarray(sd["theta"], dtype=np.float32)))
            self.m.copy_(torch.from_numpy(np.asarray(sd["m"], dtype=np.float32)))
            self.v.copy_(torch.from_numpy(np.asarray(sd["v"], dtype=np.float32)))
            self.t.copy_(torch.from_numpy(np.asarray(sd["t"], dtype=np.int32)))
            self.iteration.copy_(torch.from_numpy(np.asarray(sd["iteration"], dtype=np.int32)))
        self.ctx.sync()

    @staticmethod
    def checkpoint_path(path) -> str:
        """np.savez appends '.npz' to a name without that suffix; save and load agree on the file actually written."""
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"

    def save_checkpoint(self, path: str, **extra) -> None:
        with open(self.checkpoint_path(path), "wb") as f:
            np.savez(f, **self.state_dict(), **{k: np.asarray(v) for k, v in extra.items()})

    def load_checkpoint(self, path: str) -> dict:
        with np.load(self.checkpoint_path(path)) as z:
            sd = {k: z[k] for k in z.files}
        self.load_state_dict(sd)
        return sd

    def set_weights(self, w_hat: float, w_ind: float) -> None:
        check(lib.fbsdej_solver_set_weights(self.handle, w_hat, w_ind))

    def read_device(self, ptr: int, n: int, dtype=np.float32) -> np.ndarray:
        host = np.empty(n, dtype=dtype)
        check(lib.fbsdej_memcpy_d2h(self.ctx.handle, host.ctypes.data, ptr, host.nbytes))
        return host

    # ---- noise -----------------------------------------------------------------------------------------------
    def simulate(self, seed: int, iteration: int, B: int, path_offset: int = 0) -> None:
        check(lib.fbsdej_solver_simulate(self.handle, seed, iteration, path_offset, B))

    def set_noise(self, B: int, a=None, b=None, c=None) -> None:
        """Pricing: a = dW [N,d,B] (Merton), b = J [N,d,B], c = JMC [N,d,M]; MFG: dW0, dW, dN [N,B]."""
        ts = [None if x is None else self.ctx.to_device(x) for x in (a, b, c)]
        self._keep = ts
        check(lib.fbsdej_solver_set_noise(self.handle, B, _p(ts[0]), _p(ts[1]), _p(ts[2])))

    def set_noise_sparse_jumps(self, B: int, dW, J, jmc=None) -> None:
        """set_noise with the jump planes J [N,d,B] shipped as their non-zero entries (index, value)."""
        Jf = np.ascontiguousarray(np.asarray(J, dtype=np.float32)).reshape(-1)
        idx = np.flatnonzero(Jf).astype(np.uint32)
        ts = [None if dW is None else self.ctx.to_device(dW), self.ctx.to_device(idx.view(np.int32), dtype=torch.int32),
              self.ctx.to_device(Jf[idx]), None if jmc is None else self.ctx.to_device(jmc)]
        self._keep = ts
        check(lib.fbsdej_solver_set_noise_sparse_jumps(self.handle, B, _p(ts[0]), _p(ts[1]), _p(ts[2]), int(idx.size), _p(ts[3])))

    # ---- evaluation ------------------------------------------------------------------------------------------
    def loss(self, B: int, B_global: Optional[int] = None, traj: bool = False):
        Bg = B if B_global is None else B_global
        tx = ty = tz = None
        if traj:
            N, d = self.N, self.d
            if self.model_kind == L.MODEL_MFG:
                tx, ty = self.ctx.zeros(N + 1, 2, B), self.ctx.zeros(N + 1, 2, B)
            else:
                tx, ty, tz = self.ctx.zeros(N + 1, d, B), self.ctx.zeros(N + 1, B), self.ctx.zeros(N, d, B)
        check(lib.fbsdej_solver_loss(self.handle, _p(self.theta), B, Bg, _p(self.out), _p(tx), _p(ty), _p(tz)))
        o = self.ctx.to_host(self.out[:L.OUT_HEADER]).numpy().copy()
        if traj:
            return o, self.ctx.to_host(tx).numpy(), self.ctx.to_host(ty).numpy(), (None if tz is None else self.ctx.to_host(tz).numpy())
        return o

    def mfg_states(self, B: int) -> np.ndarray:
        """(hQ, Q, R, hS, S) of the last loss / grad call: [N+1, 5, B]."""
        out = self.ctx.zeros(self.N + 1, 5, B)
        check(lib.fbsdej_solver_mfg_states(self.handle, B, _p(out)))
        return self.ctx.to_host(out).numpy()

    def grad(self, B: int, B_global: Optional[int] = None) -> np.ndarray:
        Bg = B if B_global is None else B_global
        check(lib.fbsdej_solver_grad(self.handle, _p(self.theta), B, Bg, _p(self.out)))
        return self.ctx.to_host(self.out).numpy().copy()

    def adam_step(self, lr: float, grad: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                  beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7) -> None:
        g = self.out[L.OUT_HEADER:] if grad is None else grad
        check(lib.fbsdej_adam_step(self.ctx.handle, _p(self.theta), _p(self.m), _p(self.v), _p(g), _p(mask), self.P, lr, beta1,
                                   beta2, eps, _p(self.t)))

    def train_steps(self, seed: int, B: int, n_steps: int, lr: float, mask: Optional[torch.Tensor] = None,
                    loss_out: Optional[torch.Tensor] = None, beta1: float = 0.9, beta2: float = 0.999,
                    eps: float = 1e-7) -> None:
        check(lib.fbsdej_solver_train_steps(self.handle, _p(self.theta), _p(self.m), _p(self.v), _p(mask), _p(self.t),
                                            _p(self.iteration), seed, B, n_steps, lr, beta1, beta2, eps, _p(loss_out)))

    dp_connected = False

    # ---- data-parallel steps with the exchange inside the finishing kernel (include/fbsdej.h: fbsdej_solver_dp_*) ----------
    def dp_init(self, rank: int, world: int) -> bytes:
        h = (C.c_ubyte * 64)()
        check(lib.fbsdej_solver_dp_init(self.handle, rank, world, h))
        return bytes(h)

    def dp_buffer(self) -> int:
        p = C.c_void_p()
        check(lib.fbsdej_solver_dp_buffer(self.handle, C.byref(p)))
        return int(p.value)

    def dp_connect(self, handles: Optional[bytes] = None, raw_ptrs: Optional[list] = None) -> None:
        hb = (C.c_ubyte * len(handles)).from_buffer_copy(handles) if handles else None
        rp = (C.c_void_p * len(raw_ptrs))(*[C.c_void_p(p) if p else None for p in raw_ptrs]) if raw_ptrs else None
        check(lib.fbsdej_solver_dp_connect(self.handle, hb, rp))
        self.dp_connected = True

    def dp_check(self) -> None:
        """Syncs the stream and raises FbsdejError if a data-parallel exchange timed out (the step was not applied)."""
        check(lib.fbsdej_solver_dp_check(self.handle))

    def train_steps_dp(self, seed: int, B: int, B_global: int, path_offset: int, n_steps: int, lr: float,
                       mask: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None, beta1: float = 0.9,
                       beta2: float = 0.999, eps: float = 1e-7) -> None:
        check(lib.fbsdej_solver_train_steps_dp(self.handle, _p(self.theta), _p(self.m), _p(self.v), _p(mask), _p(self.t),
                                               _p(self.iteration), seed, B, B_global, path_offset, n_steps, lr, beta1, beta2,
                                               eps, _p(loss_out)))

    def grad_step(self, seed: int, B: int, B_global: int, path_offset: int) -> torch.Tensor:
        """simulate + forward + backward of this rank's shard; returns the device vector [4 + P] to all-reduce."""
        check(lib.fbsdej_solver_grad_step(self.handle, _p(self.theta), seed, _p(self.iteration), path_offset, B, B_global,
                                          _p(self.out)))
        return self.out

    def profile(self, seed: int, B: int, reps: int = 3) -> dict:
        """Mean device milliseconds per kernel class of one iteration (CUDA events inside the library)."""
        ms = (C.c_float * 5)()
        check(lib.fbsdej_solver_profile(self.handle, _p(self.theta), seed, B, reps, ms))
        return dict(zip(("sim_paths", "sim_compensator", "forward", "backward", "reduce"), [float(x) for x in ms]))

    def bump_iteration(self) -> None:
        check(lib.fbsdej_bump_u32(self.ctx.handle, _p(self.iteration)))

    def net_forward(self, k: int, x) -> np.ndarray:
        xt = self.ctx.to_device(x)
        y = self.ctx.empty(xt.shape[0], self.nets[k].nout)
        check(lib.fbsdej_solver_net_forward(self.handle, _p(self.theta), k, _p(xt), xt.shape[0], _p(y)))
        return self.ctx.to_host(y).numpy()

    def price(self, iStep: int, X) -> np.ndarray:
        """A(iStep, X); X [d, n] component planes (or [n] for d = 1)."""
        xt = self.ctx.to_device(np.atleast_2d(np.asarray(X, dtype=np.float32)))
        n = xt.shape[1]
        y = self.ctx.empty(n)
        check(lib.fbsdej_solver_price(self.handle, iStep, _p(xt), n, _p(y)))
        return self.ctx.to_host(y).numpy()

    def get_noise(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        check(lib.fbsdej_solver_get_noise(self.handle, *[C.byref(p) for p in ptrs]))
        return [p.value for p in ptrs]

    N: int = 0
    d: int = 1

"""Shared host logic of the drop-in solver classes: parameter packing, the outer training loop of the reference
(`train(batchSize, batchSizeVal, num_epoch, num_epochExt)`, e.g. SolversJumpDiff.py:57-73) and batch data parallelism.

Inner loop (`for epoch in range(num_epoch): trainOpt(...)`): one C-ABI call replays a CUDA graph of
simulate -> forward -> backward -> reduce -> Adam `num_epoch` times.  With torch.distributed initialised
(world_size > 1) the Monte-Carlo batch is sharded over the ranks instead and each step is
grad_step -> all_reduce([loss | grads], NCCL) -> Adam, identical on every rank.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L
from .nets import DenseNet
from .runtime import Context, NativeSolver

SEED_VALIDATION = 0x76616C69      # second Philox key word family for validation draws


def dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def shard(B: int, rank: int, world: int):
    """(offset, local count) of `rank`'s contiguous share of B paths (first B % world ranks take one more)."""
    base, rem = divmod(B, world)
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, cnt


def attach_peers(s: NativeSolver, dist, rank: int, world: int) -> bool:
    """Exchange the CUDA IPC handles of the ranks' exchange buffers (one all_gather, once per solver).  Returns False - on
    every rank - if any rank could not map its peers (no peer access between the devices): the caller then keeps the NCCL
    all-reduce loop."""
    if s.dp_connected:
        return True
    if getattr(s, "dp_unavailable", False):
        return False
    ok = 1
    try:
        h = s.dp_init(rank, world)
    except Exception:   # noqa: BLE001
        h, ok = bytes(64), 0
    mine = torch.tensor(list(h), dtype=torch.uint8, device=s.ctx.device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    if ok:
        try:
            s.dp_connect(handles=b"".join(bytes(x.cpu().tolist()) for x in parts))
        except Exception:   # noqa: BLE001
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=s.ctx.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        s.dp_connected, s.dp_unavailable = False, True
        return False
    return True


class TrainLoop:
    """Steps a NativeSolver; hides single-GPU graph replay vs. data-parallel stepping."""

    def __init__(self, native: NativeSolver, lr: float, seed: int):
        self.native, self.lr, self.seed = native, float(lr), int(seed)

    def steps(self, B: int, n: int, mask: Optional[torch.Tensor] = None) -> None:
        s = self.native
        dist, rank, world = dist_info()
        if dist is None:
            s.train_steps(self.seed, B, n, self.lr, mask=mask)
        else:
            off, cnt = shard(B, rank, world)
            if cnt == 0:
                raise ValueError(f"batch {B} is smaller than the number of ranks {world}")
            if (dist.get_backend() == "nccl" and os.environ.get("FBSDEJ_DP", "p2p") != "nccl"
                    and attach_peers(s, dist, rank, world)):
                # the step's [loss | gradient] exchange runs inside the finishing kernel over peer memory: one CUDA graph
                # per step on every rank, no collective call (FBSDEJ_DP=nccl keeps the all_reduce loop below)
                s.train_steps_dp(self.seed, cnt, B, off, n, self.lr, mask=mask)
                s.dp_check()       # syncs; raises if a peer never delivered its vector (the step is void, not NaN-poisoned)
                return
            for _ in range(n):
                out = s.grad_step(self.seed, cnt, B, off)
                with torch.cuda.stream(s.ctx.stream):
                    dist.all_reduce(out)
                s.adam_step(self.lr, mask=mask)
                s.bump_iteration()
        s.ctx.sync()

    def validation(self, B: int, draw: int) -> np.ndarray:
        """Forward-only loss on a fresh batch (every rank evaluates the same full batch)."""
        s = self.native
        s.simulate(self.seed ^ SEED_VALIDATION, draw, B)
        return s.loss(B)


class PricingSolverBase:
    """Common part of the 7 + 7 pricing solver classes.  Subclasses set SCHEME / TWO_NET / REG / Y0_NET /
    TRAIN_MULT / VAL_MULT."""

    SCHEME = L.GLOBAL
    TWO_NET = True
    REG = False
    Y0_NET = None          # 'UZ' | 'Gam' : which network object carries the trainable Y0 (Global only)
    TRAIN_MULT = 1
    VAL_MULT = 1
    M_DEFAULT = 5000       # compensator samples, hard-coded in the reference (SolversJumpDiff.py:34)

    def __init__(self, mathModel, netA: DenseNet, netB: Optional[DenseNet], lRate: float, M: Optional[int] = None,
                 seed: int = 0, ctx: Optional[Context] = None, stale_time: bool = True, tensor_cores: Optional[bool] = None):
        self.mathModel, self.netA, self.netB, self.lRate = mathModel, netA, netB, lRate
        self.M = self.M_DEFAULT if M is None else int(M)
        self.seed, self.ctx, self.stale_time = seed, ctx, stale_time
        # tcgen05 path (3xTF32 forward, bf16x3 adjoint; tests/test_tc_gpu.py): used unless told otherwise (None = automatic:
        # on when the kernels cover the scheme and the network shape - the compensator-free solvers and the jump evaluations of
        # the jump schemes at d in {1, 10})
        self.tensor_cores = tensor_cores
        self.native: Optional[NativeSolver] = None

    # ------------------------------------------------------------------------------------------------------
    def build(self) -> NativeSolver:
        if self.native is not None:
            return self.native
        mm, d = self.mathModel, self.mathModel.d
        brown = mm.kind == L.MODEL_MERTON
        self.netA.build(1 + d)
        nets = [self.netA]
        if self.TWO_NET:
            self.netB.build(1 + 2 * d)
            nets.append(self.netB)
        if self.SCHEME == L.GLOBAL:
            want = d if brown else 1
        elif self.REG:
            want = 1
        else:
            want = 1 + d if brown else 1
        if self.netA.ndimOut != want:
            raise ValueError(f"{type(self).__name__}: first network must have ndimOut={want}, got {self.netA.ndimOut}")
        if self.TWO_NET and self.netB.ndimOut != 1:
            raise ValueError("the jump network must have ndimOut=1")
        n_y0 = 1 if self.SCHEME == L.GLOBAL else 0
        M = 0 if self.REG else self.M
        spec = self.netA.spec()
        tc_ok = self.REG and spec.H <= 22 and spec.L == 2 and spec.nout == 1 and d in (1, 10)
        # jump schemes: the jump evaluations (own jump + compensator rows) on tcgen05
        jtc_ok = False
        if not self.REG and d in (1, 10):
            sb = self.netB.spec() if self.TWO_NET else spec
            jtc_ok = sb.H <= 22 and spec.H <= 23 and sb.L == 2 and spec.L == 2 and sb.activation == "tanh"
        auto = tc_ok or jtc_ok
        if self.tensor_cores and not auto:
            raise ValueError(f"{type(self).__name__}: tensor_cores=True, but the tcgen05 kernels cover two hidden layers of width <= 22 "
                             f"(tanh for the jump schemes) at d in (1, 10); got d={d}, first network {spec}")
        use_tc = auto if self.tensor_cores is None else bool(self.tensor_cores)
        self.native = mm.make_solver(self.SCHEME, [n.spec() for n in nets], n_y0, M, ctx=self.ctx,
                                     stale_time=self.stale_time, tensor_cores=use_tc)
        self.push_params()
        return self.native

    def _y0_holder(self):
        return self.netA if self.Y0_NET == "UZ" else self.netB

    def push_params(self) -> None:
        parts = [self.netA.params] + ([self.netB.params] if self.TWO_NET else [])
        if self.SCHEME == L.GLOBAL:
            holder = self._y0_holder()
            if not hasattr(holder, "Y0"):
                raise ValueError("Global solver: the network carrying Y0 must be built with bY0=1")
            parts.append(np.array([holder.Y0.numpy()], dtype=np.float32))
        self.native.set_theta(np.concatenate(parts))

    def pull_params(self) -> None:
        th = self.native.get_theta()
        offs = self.native.offsets
        self.netA.params = th[offs[0]:offs[0] + self.netA.params.size].copy()
        if self.TWO_NET:
            self.netB.params = th[offs[1]:offs[1] + self.netB.params.size].copy()
        if self.SCHEME == L.GLOBAL:
            self._y0_holder().Y0.assign(th[self.native.y0_offset])

    def current_Y0(self) -> float:
        s = self.native
        if self.SCHEME == L.GLOBAL:
            return float(s.get_theta()[s.y0_offset])
        # mean of net(0, x0) over 10^5 identical rows (SolversJumpDiff.py:140-141) == one row
        x = np.array([[0.0] + [self.mathModel.x0] * self.mathModel.d], dtype=np.float32)
        return float(s.net_forward(0, x)[0, 0])

    # ------------------------------------------------------------------------------------------------------
    def train(self, batchSize, batchSizeVal, num_epoch, num_epochExt):
        s = self.build()
        loop = TrainLoop(s, self.lRate, self.seed)
        resumed, self._resumed = getattr(self, "_resumed", False), False
        if not resumed:                    # a fresh optimizer per train() call (SolversJumpDiff.py:55) - unless load() just
            s.reset_optimizer()            # restored Adam's m, v, t and the Philox iteration: then training continues bit-exactly
            self.listY0: List[float] = []
            self.lossList: List[float] = []
        self.duration = 0
        self.durationList: List[float] = []
        B, Bval = self.TRAIN_MULT * batchSize, self.VAL_MULT * batchSizeVal
        for iout in range(num_epochExt):
            start_time = time.time()
            loop.steps(B, num_epoch)
            self.duration += time.time() - start_time
            objError = float(loop.validation(Bval, iout)[0])
            Y0 = self.current_Y0()
            _, rank, _ = dist_info()
            if rank == 0:
                print(" Error", objError, " elapsed time %5.3f s" % self.duration, "Y0 sofar ", Y0, 'epoch', iout)
            self.listY0.append(np.float32(Y0))
            self.lossList.append(objError)
            self.durationList.append(self.duration)
        self.pull_params()
        return self._result()

    # checkpoint / resume
    def save(self, path: str) -> None:
        self.build().save_checkpoint(path, seed=self.seed, listY0=getattr(self, "listY0", []), lossList=getattr(self, "lossList", []))

    def load(self, path: str) -> None:
        sd = self.build().load_checkpoint(path)
        self.seed = int(sd.get("seed", self.seed))
        self.listY0 = [np.float32(x) for x in sd.get("listY0", [])]
        self.lossList = [float(x) for x in sd.get("lossList", [])]
        self._resumed = True               # the next train() keeps the restored optimizer state
        self.pull_params()

    def _result(self):
        return self.listY0, self.duration

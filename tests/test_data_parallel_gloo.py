"""CPU, world_size 2, gloo: the data-parallel convention of SURVEY 8e.  Each rank evaluates its shard of the batch
with the GLOBAL 1/B weight; one all-reduce(sum) of [loss | grads] must reproduce the single-process loss and gradient,
and the Keras-form Adam step that follows is then identical on every rank.  (The per-rank evaluation is done by the
oracle here; on the GPU the same vector comes from fbsdej_solver_grad_step, covered by tests/test_dp_gpu.py.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _worker(rank, world, port, q):
    import helpers as H
    from oracle import MertonOracle, KerasAdam, pricing_loss
    from deepfbsdejsolvers_b200.solver_base import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    B, scheme = 11, "SumLocalReg"
    om = MertonOracle(aLin=H.ALIN, limit=30, d=1, **dict(H.MERTON, N=8))
    layout = H.pricing_layout("merton", scheme, 1)
    theta = H.random_theta(layout, 1)
    noise = H.merton_noise(om, B, 0, seed=3, with_jmc=False)
    off, cnt = shard(B, rank, world)
    th = torch.tensor(theta, requires_grad=True)
    local = {k: v[:, off:off + cnt] for k, v in noise.items()}
    part = pricing_loss(om, scheme, layout, th, local, cnt) * (cnt / B)      # this rank's share of the global mean
    part.backward()
    vec = torch.cat([part.detach().reshape(1), torch.zeros(3), th.grad])
    dist.all_reduce(vec)
    new = torch.tensor(theta.copy())
    KerasAdam(layout.total, 3e-4).step(new, vec[4:])
    if rank == 0:
        full = torch.tensor(theta, requires_grad=True)
        loss = pricing_loss(om, scheme, layout, full, noise, B)
        loss.backward()
        ref = torch.tensor(theta.copy())
        KerasAdam(layout.total, 3e-4).step(ref, full.grad)
        q.put((float(loss), float(vec[0]), float((vec[4:] - full.grad).abs().max() / full.grad.abs().max()),
               float((new - ref).abs().max())))
    gathered = [torch.zeros_like(new) for _ in range(world)]
    dist.all_gather(gathered, new)
    assert all(torch.equal(gathered[0], g) for g in gathered)      # replicas stay bit-identical
    dist.destroy_process_group()


def test_sharded_sum_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    loss, loss_dp, gerr, therr = q.get(timeout=10)
    assert abs(loss - loss_dp) <= 1e-6 * abs(loss)
    assert gerr < 1e-5 and therr < 1e-6

"""world_size-2 gloo tests of the host-side data-parallel logic (CPU): sharding, flat-bucket all-reduce + 1/world
scaling == the mean gradient a single process computes on the union of the shards (for a BatchNorm-free model, where
sharding does not change the maths), and replica broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wsl4mis_b200 import ddp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert ddp.env_world() == (rank, rank, world)
        torch.manual_seed(0)
        w = torch.randn(16, 1, 3, 3)                     # a BN-free "model": conv + mean-square loss
        x = torch.randn(6, 1, 8, 8)
        lo, hi = ddp.shard_range(6, rank, world)
        wr = w.clone().requires_grad_(True)
        loss = torch.nn.functional.conv2d(x[lo:hi], wr, padding=1).pow(2).mean()
        (g,) = torch.autograd.grad(loss, wr)
        bucket = g.flatten().clone()
        ddp.allreduce_flat(bucket)
        mean_grad = bucket / world                       # what wsl_sgd_step(grad_scale=1/world) consumes
        wf = w.clone().requires_grad_(True)
        full = torch.nn.functional.conv2d(x, wf, padding=1).pow(2).mean()
        (gf,) = torch.autograd.grad(full, wf)
        ok = torch.allclose(mean_grad, gf.flatten(), atol=1e-6)
        p = torch.full((5,), float(rank))
        ddp.broadcast_flat(p, 0)
        ok = ok and bool((p == 0).all())
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    for n in (1, 7, 64, 128):
        for world in (1, 2, 3, 8):
            spans = [ddp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_allreduce_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}

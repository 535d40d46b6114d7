"""Batch-sharded data parallelism for the hot path: one process per GPU, weights replicated, ONE all-reduce of the
flat fp32 gradient bucket per step (the reference is single-process; this is an addition, not a port).

Host-side helpers only; they work on any backend (NCCL on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process -> (0, 0, 1))."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a global batch of n_items images for `rank` (images are independent units; the
    remainder goes to the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat(bucket: torch.Tensor, group=None):
    """Sum the flat gradient bucket over all ranks in place (the SGD kernel applies the 1/world factor)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket


def broadcast_flat(params: torch.Tensor, src=0, group=None):
    """Make every rank start from rank `src`'s flat parameter buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(params, src=src, group=group)
    return params

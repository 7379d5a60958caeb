"""Multi-GPU plumbing of the training steps: one process per GPU, minibatch images
sharded over ranks, ONE all-reduce (sum) per step over a flat buffer that holds every parameter-shaped
statistic (clamped - unclamped) plus the loss, then the identical SGD update on every rank.

The reference has no distributed layer (its only parallelism is a process pool of sampler calls,
src/model/faster_dqbm.py:98-111,578-596); the sum-then-divide-by-the-global-batch order matches
src/train/train.py:101-112 and src/model/faster_dqbm.py:1042-1049.  The functions here work on tensors
of any device, so the same code runs under NCCL on GPUs and under gloo in the CPU tests (the update itself is the
native K9 `qbm_sgd_apply` on the flat parameter buffer).
"""
from __future__ import annotations

import torch


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of `total` units owned by `rank` (the first total % world ranks own one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} for world size {world}")
    base, extra = divmod(int(total), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_reduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over the ranks of `group`; a no-op without a process group."""
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat

"""Multi-GPU plumbing of the evaluation path: independent replicas, batch-sharded.

The forward pass has no cross-sample arithmetic (reference ``src/models/fortitran.py:184-233``), so the only
collectives of the path are the ones that follow it in the evaluator: an all-gather of the
complex64 estimates and a SUM all-reduce of the error / power accumulators
(reference ``src/main/trainer.py:338-345``).  One process per GPU (``torchrun``), NCCL on GPUs, gloo in the CPU
tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard ``[lo, hi)`` of ``batch`` samples for ``rank`` (first ``batch % world`` ranks get one more)."""
    if batch < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad shard request batch={batch} rank={rank} world={world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_estimates(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equal-sized complex64 shards ``[b, scs, sym]`` into ``[world * b, scs, sym]`` (rank order)."""
    if not torch.is_complex(local):
        raise TypeError("estimates must be complex")
    world = dist.get_world_size(group)
    local = local.contiguous()
    out = torch.empty((world * local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(torch.view_as_real(out), torch.view_as_real(local), group=group)
    return out


def reduce_error_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of ``[sum |est - truth|^2, sum |truth|^2, ...]`` accumulators (fp64), in place."""
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def mse_db_from_sums(sums: torch.Tensor, count: int) -> float:
    """The reference's reported metric, ``to_db(2 * MSELoss(cat(re, im)))`` == 10 log10(sum|e|^2 / count)."""
    return float(10.0 * torch.log10(sums[0] / count))

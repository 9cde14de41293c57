"""Multi-GPU plumbing of the TimesBlock path: one process per GPU, batch (window / series) sharded.

Every stage of the path is per-window except one reduction: the batch mean of the channel-median
amplitude spectrum inside the shared period search (reference timesnet.py:112).  With the batch
sharded over ranks each rank sums its own windows (``ftn_spectrum`` -> ``amp_sum[F]``), the sums are
all-reduced -- ``L/2 + 1`` floats, the only collective of the path -- and every rank then runs the
identical deterministic top-k / grouping tail on the identical vector, so all ranks fold with the
same periods.  The functions below are device agnostic (NCCL on the B200
box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_batch(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[lo, hi)`` slice of the batch owned by ``rank`` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(batch), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def resolve_group(process_group) -> Tuple[Optional[object], int]:
    """``process_group`` semantics of ``FFTPeriodSelector``: ``None`` = default group when
    torch.distributed is initialised, ``False`` = never reduce.  Returns ``(group, world_size)``."""
    import torch.distributed as dist
    if process_group is False or not dist.is_available() or not dist.is_initialized():
        return None, 1
    return process_group, dist.get_world_size(process_group)


def reduce_spectrum_sum(amp_sum: torch.Tensor, local_batch: int, process_group=None) -> Tuple[torch.Tensor, int]:
    """All-reduce (SUM) the batch-summed spectrum in place and return it with the global window count.

    Ranks may hold different numbers of windows (ragged last shard), so the count is reduced too --
    it rides in the same message as one extra float-exact integer slot (counts < 2**24)."""
    import torch.distributed as dist
    group, world = resolve_group(process_group)
    if world == 1:
        return amp_sum, int(local_batch)
    msg = torch.empty(amp_sum.numel() + 1, dtype=torch.float32, device=amp_sum.device)
    msg[:-1] = amp_sum.reshape(-1).to(torch.float32)
    msg[-1] = float(local_batch)
    dist.all_reduce(msg, op=dist.ReduceOp.SUM, group=group)
    amp_sum.copy_(msg[:-1].reshape(amp_sum.shape))
    return amp_sum, int(round(float(msg[-1].item())))

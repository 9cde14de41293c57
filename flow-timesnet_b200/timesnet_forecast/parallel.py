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
    """``process_group`` semantics of ``FFTPeriodSelector``: ``False`` (the default) = never
    reduce, ``None`` = default group when torch.distributed is initialised, else the given group.  Returns ``(group, world_size)``."""
    import torch.distributed as dist
    if process_group is False or not dist.is_available() or not dist.is_initialized():
        return None, 1
    return process_group, dist.get_world_size(process_group)


def share_period_search(model_or_selector, process_group=None, transport: str = "auto"):
    """Opt in to the SHARED period search of SURVEY.md section 8e: the batch-summed amplitude spectrum is summed over
    ``process_group`` (``None`` = the default group) so that every rank folds with the same periods -- the periods the
    single-process reference would pick on the concatenated batch.  Without this call every rank selects from its local
    batch, which is what a reference checkout running DDP does.

    ``transport``:
      * ``"peer"`` -- NVLink peer mailbox (csrc/peer.cuh): every rank's selection kernel stores its L/2 + 2 partial sums
        into the peers' memory and adds the rows up in rank order; no NCCL call, and search + block stay one library
        call (``ftn_timesblock_forward``).  Needs CUDA, one process per GPU on one node, world <= 16.
      * ``"nccl"`` -- ``torch.distributed.all_reduce`` on the group's backend (NCCL on GPUs, gloo in the CPU tests).
      * ``"auto"`` -- ``"peer"`` when every rank can set it up, else ``"nccl"``.

    Accepts a ``TimesNet`` (its ``period_selector``), a ``TimesBlock`` or an ``FFTPeriodSelector``.  Every rank of the
    group must then call ``forward`` the same number of times (the search contains a collective)."""
    import torch.distributed as dist
    sel = getattr(model_or_selector, "period_selector", model_or_selector)
    if not hasattr(sel, "process_group"):
        raise TypeError("share_period_search expects a TimesNet, a TimesBlock or an FFTPeriodSelector")
    if transport not in ("auto", "peer", "nccl"):
        raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
    sel.process_group = process_group
    sel.peer_comm = None
    group, world = resolve_group(process_group)
    if world > 1 and transport in ("auto", "peer"):
        comm, err = None, None
        try:
            if not torch.cuda.is_available() or world > 16:
                raise RuntimeError("the peer mailbox needs CUDA and at most 16 ranks")
            from . import _native as nv

            def gather(mine: bytes):
                out = [None] * world
                dist.all_gather_object(out, mine, group=group)
                return out

            comm = nv.PeerComm(dist.get_rank(group), world, gather)
        except Exception as exc:                                    # noqa: BLE001 -- reported or re-raised below
            err = exc
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)       # all ranks agree on the transport
        if all(oks):
            sel.peer_comm = comm
        else:
            if comm is not None:
                comm.close()
            if transport == "peer":
                raise RuntimeError(f"NVLink peer mailbox could not be set up on every rank: {err}")
    return sel


def local_period_search(model_or_selector):
    """Undo ``share_period_search``: rank-local period selection (the default)."""
    sel = getattr(model_or_selector, "period_selector", model_or_selector)
    sel.process_group = False
    if getattr(sel, "peer_comm", None) is not None:
        sel.peer_comm.close()
    sel.peer_comm = None
    return sel


def reduce_spectrum_sum(msg: torch.Tensor, process_group=None) -> torch.Tensor:
    """The one collective of the path: all-reduce (SUM), in place, the ``[F + 1]`` fp32 message ``ftn_spectrum``
    writes -- ``F`` batch-summed channel-median amplitudes followed by the number of windows summed.  Ranks may hold
    different numbers of windows (ragged last shard, even zero); the count rides in the same message as one
    float-exact integer slot (counts < 2**24), so the selection tail divides by the GLOBAL batch without a host
    round trip.  Called by ``FFTPeriodSelector.search`` (NCCL on the B200 box) and by the CPU tests (gloo)."""
    import torch.distributed as dist
    group, world = resolve_group(process_group)
    if world > 1:
        dist.all_reduce(msg, op=dist.ReduceOp.SUM, group=group)
    return msg

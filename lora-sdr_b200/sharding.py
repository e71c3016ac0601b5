"""Host-side multi-GPU logic: how symbol windows / channels are split across ranks
and how rank 0's tables reach the others.  

The hot path has no exchange step: windows (and channels) are independent, so each
rank demodulates a contiguous block of the window index space and no collective
runs on the data path.  The only collective is the one-time broadcast of the
dechirp / fine-tune / twiddle tables from rank 0 (O(KB..MB)); results stay on the
rank that produced them unless the caller gathers them.

Everything here is backend-agnostic (`torch.distributed` with "nccl" on GPUs,
"gloo" in the CPU tests).
"""
from __future__ import annotations

from typing import Sequence


def window_range(n_windows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition: rank r of R gets windows [r*W/R, (r+1)*W/R)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if n_windows < 0:
        raise ValueError("n_windows must be >= 0")
    return n_windows * rank // world, n_windows * (rank + 1) // world


def channel_range(n_channels: int, rank: int, world: int) -> tuple[int, int]:
    """Whole channels per rank (BASELINE.json config 4: 256 channels x SF10 over 8 GPUs = 32 each);
    keeps every channel's stream contiguous on one GPU."""
    return window_range(n_channels, rank, world)


def broadcast_tables(tables: Sequence, src: int = 0, group=None) -> None:
    """In-place broadcast of the table tensors (float32 views) from `src` to every rank."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tables:
        dist.broadcast(t, src=src, group=group)


def gather_results(local, group=None):
    """all_gather of per-rank result tensors (same length on every rank) -> list in rank order."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local]
    out = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, local, group=group)
    return out

"""Clip sharding across ranks: the propagation path has no cross-clip state
(eval_interactive_davis.py:76-83 builds a fresh processor per sequence), so a batch of clips is
dealt round-robin to the ranks of one node, weights are replicated per rank, and the only
collectives are a barrier plus tiny reductions/gathers of timings and per-clip checksums
(NCCL on GPUs, gloo in the CPU tests).  No per-frame collective exists on this path."""
from __future__ import annotations

from typing import List, Sequence


def clips_of_rank(num_clips: int, rank: int, world: int) -> List[int]:
    """Clip c -> rank c % world."""
    assert 0 <= rank < world
    return list(range(rank, num_clips, world))


def owner_of_clip(clip: int, world: int) -> int:
    return clip % world


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank device time (the bench contract: max over ranks, never wall clock)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_clip_results(local: Sequence, num_clips: int) -> List:
    """All-gather per-rank result lists [(clip_index, payload), ...] into one list ordered by clip
    index on every rank; checks that every clip was produced exactly once."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, list(local))
    else:
        parts = [list(local)]
    merged = {}
    for part in parts:
        for idx, payload in part:
            assert idx not in merged, f"clip {idx} produced twice"
            merged[idx] = payload
    assert sorted(merged) == list(range(num_clips)), "missing clips: %r" % (sorted(set(range(num_clips)) - set(merged)),)
    return [merged[i] for i in range(num_clips)]

"""Drop-in soft aggregation (reference: model/aggregate.py).  CUDA tensors only."""
from __future__ import annotations

import torch

from . import ops


def aggregate_wbg(prob: torch.Tensor, keep_bg: bool = False, hard: bool = False) -> torch.Tensor:
    """aggregate.py:22-37: bg = prod(1-p); clamp; logit (x1000 if hard); softmax over K+1."""
    return ops.aggregate_wbg(prob.detach().float().contiguous(), keep_bg=keep_bg, hard=hard)


def aggregate_sbg(prob: torch.Tensor, keep_bg: bool = False, hard: bool = False) -> torch.Tensor:
    """aggregate.py:4-20: constant 0.5 background instead of prod(1-p)."""
    return ops.aggregate_wbg(prob.detach().float().contiguous(), keep_bg=keep_bg, hard=hard, const_bg=True)

"""Multi-GPU plumbing: frames are independent, so a batch is block-partitioned over the ranks with no
data-path collective; one all-gather of the fixed-size outputs collects the paths.

One process per GPU (`torchrun`), `torch.distributed` with the nccl backend on GPUs; the same code runs on
the gloo backend with CPU tensors, which is how tests/test_dist_gloo.py covers it without a GPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of ceil(B / G) frames per rank; trailing ranks may get fewer (or none)."""
    per = (n_frames + world_size - 1) // world_size
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def all_gather_frames(local: torch.Tensor, n_frames: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Gather per-frame outputs ([n_local, ...], rank-contiguous blocks as in shard_bounds) into [n_frames, ...].

    Every rank contributes a block padded to ceil(B / G) rows so that one equal-count all-gather suffices."""
    world = dist.get_world_size(group)
    per = (n_frames + world - 1) // world
    tail = local.shape[1:]
    if local.shape[0] != per:
        padded = local.new_zeros((per, *tail))
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    out = local.new_empty((world * per, *tail))
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_frames]

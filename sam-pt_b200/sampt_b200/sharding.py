"""Frame sharding of a clip across ranks: rank r owns frames {f : f mod G == r} (round-robin keeps every
8-frame tracker window spread over all ranks and balances the per-frame encoder work).  The one exchange step of the path
is an all-gather of per-frame tracker feature maps; this module holds the index arithmetic + the collective wrapper so it
can be tested with gloo on CPU (world_size 2) and run with NCCL on the B200s."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def owned_frames(T: int, rank: int, world: int) -> List[int]:
    return list(range(rank, T, world))


def padded_count(T: int, world: int) -> int:
    """frames per rank after padding so every rank contributes the same number of rows to the all-gather."""
    return (T + world - 1) // world


def allgather_frames(local: torch.Tensor, T: int, group=None) -> torch.Tensor:
    """local: (n_owned, ...) features of this rank's frames (in increasing frame order) -> (T, ...) in frame order on
    every rank.  One collective (NCCL all_gather_into_tensor over NVLink on GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = padded_count(T, world)
    assert local.shape[0] == len(owned_frames(T, rank, world))
    if local.shape[0] < n:  # pad the short ranks (T not divisible by world)
        pad = torch.zeros((n - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    out = torch.empty((world * n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    # row (r*n + i) holds frame r + i*world  ->  reorder to frame order, dropping the padding
    idx = torch.tensor([(f % world) * n + f // world for f in range(T)], device=local.device)
    return out.index_select(0, idx)


def scatter_rows_by_frame(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """inverse view: the rows of a (T, ...) tensor this rank owns."""
    return full[rank::world]

"""Camera-sharded data parallelism: every rank holds all Gaussian parameters, renders
its own camera, and the flat per-Gaussian gradient arena is summed across ranks with ONE all-reduce.
No pixel / Gaussian partitioning: "the render is replicated (per camera)" (BASELINE.json north_star).
The reference itself has no distributed code."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def camera_for_rank(step: int, rank: int, world_size: int, num_cameras: int) -> int:
    """Rank r of g renders camera (step*g + r) mod #cameras: g distinct cameras per step."""
    return (step * world_size + rank) % num_cameras


def allreduce_gradients(arena: torch.Tensor, average: bool = False, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place SUM (or mean) of the flat gradient arena produced by the project/SH/compose backward.
    Every parameter gradient is a view of the arena, so this is the only collective of a training step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arena
    dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=group)
    if average:
        arena.div_(dist.get_world_size(group))
    return arena


def allreduce_densification_stats(xys_grad_norm: torch.Tensor, vis_counts: torch.Tensor, max_2dsize: torch.Tensor,
                                  group: Optional[dist.ProcessGroup] = None) -> None:
    """The small per-Gaussian statistics every replica needs to take identical split/cull decisions
    (street_gaussians_ns/sgn_splatfacto.py:520-541): SUM, SUM, MAX."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(xys_grad_norm, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(vis_counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_2dsize, op=dist.ReduceOp.MAX, group=group)

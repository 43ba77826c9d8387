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


def chunk_bounds(total: int, chunks: int, align: int = 1024) -> list:
    """``chunks`` contiguous ranges of the arena, ends aligned to ``align`` floats (a multiple of 4: 16-byte slices)."""
    assert align % 4 == 0 and chunks >= 1
    edges = [min(total, (total * k // chunks + align - 1) // align * align) for k in range(chunks + 1)]
    edges[0], edges[-1] = 0, total
    return [(lo, hi) for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]


def allreduce_and_step(arena: torch.Tensor, optimizer, present, chunks: int = 4,
                       group: Optional[dist.ProcessGroup] = None) -> None:
    """Gradient all-reduce (mean over the replicas) and the fused Adam step, pipelined over ``chunks`` ranges of the
    arena: all ranges' all-reduces are enqueued at once (they run in order on the communicator's stream), and Adam is
    launched on range k as soon as ITS all-reduce has finished, while range k+1 is still on the wire -- the optimizer
    (28 B per element through HBM) hides behind the collective instead of following it.  Needs the arena in the
    optimizer's layout (``SceneGraphConfig.full_gradient_arena``).  ``optimizer``: a FusedAdam; ``present``: the
    sub-models that have a gradient on some replica (None = all)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    optimizer.step_count += 1
    tab = optimizer.step_table(present, full_layout=True)
    if world == 1:
        optimizer.launch(tab, arena)
        return
    bounds = chunk_bounds(optimizer.arena_elems, chunks)
    avg = dist.get_backend(group) == "nccl"  # NCCL averages inside the collective; gloo (CPU tests) sums
    works = [dist.all_reduce(arena[lo:hi], op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=group, async_op=True)
             for lo, hi in bounds]
    for (lo, hi), work in zip(bounds, works):
        work.wait()  # NCCL: the current stream waits for this range's collective only
        if not avg:
            arena[lo:hi].div_(world)
        optimizer.launch(optimizer.rows_in_range(tab, lo, hi), arena)

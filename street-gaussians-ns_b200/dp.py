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


# --------------------------------------------------------------------------------------------------------------------
# The exchange as this library's own kernel over symmetric memory (csrc/collective.cu)
# --------------------------------------------------------------------------------------------------------------------
def plan_ranges(seg_counts, seg_arena_offsets, seg_widths, num_ranges: int, chunk_rows: int = 128):
    """Cuts the project backward of a frame into ``num_ranges`` chunk ranges and names the arena slices each one writes.

    ``seg_counts[s]``: rows of segment s (frame order); ``seg_arena_offsets[s][k]`` / ``seg_widths[s][k]``: where tensor k of
    segment s starts in the arena (floats) and its floats per row.  The first (largest) segment -- the background -- is cut
    into row ranges at chunk boundaries; every other segment stays whole, and runs of whole segments whose arena regions are
    adjacent collapse into one slice.  Returns [(chunk_begin, chunk_end, [slice, ...]), ...] with slice = (offset, length) or,
    for the first segment, (offset, length, floats per row, first row, rows); lengths are rounded up to 4 floats (the arena pads
    every tensor to 16 bytes).  Plain integer arithmetic (CPU-tested)."""
    def pad4(x):
        return (x + 3) // 4 * 4
    chunks = [(n + chunk_rows - 1) // chunk_rows for n in seg_counts]
    chunk0 = [sum(chunks[:i]) for i in range(len(chunks))]
    out = []
    head = max(1, num_ranges - 1) if len(seg_counts) > 1 else max(1, num_ranges)
    per = (chunks[0] + head - 1) // head if chunks[0] else 0
    c = 0
    while c < chunks[0]:
        c1 = min(chunks[0], c + per)
        r0, r1 = c * chunk_rows, min(seg_counts[0], c1 * chunk_rows)
        # (offset, length, floats per row, first row, rows): the row description lets the exchange skip rows no replica saw
        sl = [(int(seg_arena_offsets[0][k] + r0 * seg_widths[0][k]), int(pad4((r1 - r0) * seg_widths[0][k])), int(seg_widths[0][k]), int(r0),
               int(r1 - r0)) for k in range(6)]
        out.append((chunk0[0] + c, chunk0[0] + c1, [x for x in sl if x[1] > 0]))
        c = c1
    if len(seg_counts) > 1:
        runs = []
        for s in range(1, len(seg_counts)):
            for k in range(6):
                o, ln = int(seg_arena_offsets[s][k]), int(pad4(seg_counts[s] * seg_widths[s][k]))
                if ln == 0:
                    continue
                if runs and runs[-1][0] + runs[-1][1] == o:
                    runs[-1][1] += ln
                else:
                    runs.append([o, ln])
        out.append((chunk0[1], chunk0[-1] + chunks[-1], [tuple(r) for r in runs]))
    return out


def frame_arena_layout(frame):
    """(row counts, per-tensor arena offsets, floats per row, total floats) of the gradient arena project_bwd writes for a
    frame: segment-major, six tensors per segment, every tensor padded to 16 bytes (raster.arena_layout)."""
    counts, widths, offs, cur = [], [], [], 0
    for seg in frame.segments:
        p = seg.params
        n = int(p.means.shape[0])
        w = [3, 3, 4, 3 * int(p.features_dc.shape[1]), 3 * int(p.features_rest.shape[1]), 1]
        row = []
        for k in range(6):
            row.append(cur)
            cur += (n * w[k] + 3) // 4 * 4
        counts.append(n)
        widths.append(w)
        offs.append(row)
    return counts, offs, widths, cur


class SymmetricExchange:
    """The gradient arena in symmetric memory + the exchange kernels over it.

    ``arena`` is the tensor the project backward writes (and Adam reads); ``all_reduce(slices)`` sums (or averages) the listed
    slices over the ranks IN PLACE with ``sgn_allreduce_sym`` between two device-side barriers, on ``stream`` (default: the
    current one).  With ``comm_stream`` the exchange of range k runs next to the production of range k+1:
    ``begin()`` ... ``after_range(k, slices)`` ... ``wait_range(k)``.
    Needs CUDA peer access between the ranks' GPUs (NVLink); the multicast path additionally needs an NVSwitch fabric --
    without it the kernel pulls from / pushes to the peers' arenas directly.  ``use_multicast``: True / False / "auto"
    (time both paths once on this arena and keep the faster)."""

    def __init__(self, numel: int, device, group: Optional[dist.ProcessGroup] = None, use_multicast="auto", flag_rows: int = 0):
        """``flag_rows`` > 0: room for that many per-row visibility flags behind the arena (same symmetric allocation), so that
        the exchange can skip the rows no replica saw (``publish_visible`` / ``all_reduce(..., skip_unseen=True)``)."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._C, self._lib = C, _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.numel = int((numel + 3) // 4 * 4)
        self.flag_rows = int(flag_rows)
        flag_floats = (self.flag_rows + 15) // 16 * 4
        self._storage = symm.empty(self.numel + flag_floats, dtype=torch.float32, device=device)
        self._storage.zero_()
        self.arena = self._storage[:self.numel]
        self.flags = self._storage[self.numel:].view(torch.uint8)[:self.flag_rows] if self.flag_rows else None
        self.union = torch.zeros(max(self.flag_rows, 1), dtype=torch.uint8, device=device)
        self._union_fresh = False
        self.hdl = symm.rendezvous(self._storage, self.group)
        mc = 0
        if use_multicast:
            try:
                mc = int(self.hdl.multicast_ptr or 0)
            except Exception:
                mc = 0
        self.multicast_ptr = mc
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        self.comm_stream = torch.cuda.Stream(device=device)
        self._done = {}
        self._open = None
        self.tuned = None
        torch.cuda.synchronize(device)
        dist.barrier(self.group)
        if mc and use_multicast == "auto":
            self._autotune(device)

    def _autotune(self, device) -> None:
        """Which path is faster on THIS box and world size is measured, not assumed: at two GPUs the peer loads / stores beat
        the in-switch reduction (0.53 vs 0.85 ms for 330 MB on a B200 pair, profiles/), with more replicas the multicast path
        moves 1/(2 - 2/g) of the bytes.  Three exchanges of the (zeroed) arena per mode, max over ranks, every rank takes the
        same decision."""
        mc = self.multicast_ptr
        times = []
        for mode_mc in (mc, 0):
            self.multicast_ptr = mode_mc
            self.all_reduce()
            torch.cuda.synchronize(device)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                self.all_reduce()
            b.record()
            torch.cuda.synchronize(device)
            t = torch.tensor([a.elapsed_time(b) / 3], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            times.append(float(t.item()))
        self.multicast_ptr = mc if times[0] <= times[1] else 0
        self.tuned = {"multimem_ms": round(times[0], 4), "peer_ms": round(times[1], 4)}
        self.arena.zero_()

    @property
    def mode(self) -> str:
        return "multimem (in-switch reduction)" if self.multicast_ptr else "peer loads/stores"

    def publish_visible(self, radii: torch.Tensor, rows: Optional[int] = None) -> None:
        """This rank's per-row visibility (radii > 0, the frame's first ``flag_rows`` rows -- the background, whose rows mean the
        same Gaussian on every replica) into the symmetric flags; call on the stream that produced ``radii``, before the
        backward.  The next exchange with ``skip_unseen`` ORs the replicas' flags once and skips the rows nobody saw."""
        rows = self.flag_rows if rows is None else int(rows)  # the allocation may have room for more rows than the model has now
        assert self.flags is not None and radii.dtype == torch.int32 and 0 <= rows <= min(self.flag_rows, radii.shape[0])
        C = self._C
        self._lib.check(self._lib.load().sgn_visible_flags(C.c_void_p(radii.data_ptr()), rows, C.c_void_p(self.flags.data_ptr()),
                                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sgn_visible_flags")
        self._union_fresh = False
        self._union_rows = rows

    def _launch(self, slices, scale: float, max_ctas: int = 0, skip_unseen: bool = False):
        C = self._C
        n = len(slices)
        if n == 0:
            return
        assert n <= self._lib.AR_MAX_SLICES, n
        off = (C.c_int64 * n)(*[int(sl[0]) for sl in slices])
        ln = (C.c_int64 * n)(*[int(sl[1]) for sl in slices])
        L = self._lib.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        widths = row0 = rows = union = None
        if skip_unseen and self.flags is not None and any(len(sl) == 5 for sl in slices):
            if not self._union_fresh:  # once per step, after the first barrier: every replica's flags are in place
                self._lib.check(L.sgn_visible_union(C.c_void_p(self.peers_dev), 4 * self.numel, self.world, getattr(self, "_union_rows", self.flag_rows),
                                                    C.c_void_p(self.union.data_ptr()), stream), "sgn_visible_union")
                self._union_fresh = True
            widths = (C.c_int32 * n)(*[int(sl[2]) if len(sl) == 5 else 0 for sl in slices])
            row0 = (C.c_int64 * n)(*[int(sl[3]) if len(sl) == 5 else 0 for sl in slices])
            rows = (C.c_int64 * n)(*[int(sl[4]) if len(sl) == 5 else 0 for sl in slices])
            union = C.c_void_p(self.union.data_ptr())
        self._lib.check(L.sgn_allreduce_sym(C.c_void_p(self.arena.data_ptr()), C.c_void_p(self.multicast_ptr or None),
                                            C.c_void_p(self.peers_dev), self.rank, self.world, n, off, ln, widths, row0, rows, union,
                                            C.c_float(scale), max_ctas, stream), "sgn_allreduce_sym")

    def all_reduce(self, slices=None, average: bool = False, max_ctas: int = 0, skip_unseen: bool = False):
        """In place, on the current stream.  ``slices``: [(offset, length)] in floats (multiples of 4), or with a row
        description (offset, length, floats per row, first row, rows) for ``skip_unseen``; None = the whole arena."""
        if slices is None:
            slices = [(0, self.numel)]
        self.hdl.barrier(channel=0)   # every replica has written these slices (and published its visibility flags)
        self._launch(slices, 1.0 / self.world if average else 1.0, max_ctas, skip_unseen)
        self.hdl.barrier(channel=1)   # every part has been pushed to every replica

    # ---- range by range, on the communication stream ------------------------------------------------------------
    # Barriers: range k's exchange starts behind a barrier ("every replica has written range k").  On each rank that barrier
    # also follows range k-1's kernel in stream order, so passing it means every replica has finished PUSHING range k-1: the
    # trailing barrier is only needed once, after the last range (K + 1 barriers per step instead of 2 K).
    def after_range(self, k: int, slices, average: bool = False, max_ctas: int = 0, skip_unseen: bool = False):
        """Call right after the launch that PRODUCES range k was enqueued on the current stream."""
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self.hdl.barrier(channel=0)
            if self._open is not None:  # the previous range has arrived everywhere
                done = torch.cuda.Event()
                done.record()
                self._done[self._open] = done
            self._launch(slices, 1.0 / self.world if average else 1.0, max_ctas, skip_unseen)
            self._open = k

    def _close(self):
        if self._open is None:
            return
        with torch.cuda.stream(self.comm_stream):
            self.hdl.barrier(channel=1)
            done = torch.cuda.Event()
            done.record()
            self._done[self._open] = done
        self._open = None

    def wait_range(self, k: int):
        """The current stream waits until range k has been exchanged (closes the sequence when k is the last range issued)."""
        if k not in self._done:
            assert k == self._open, (k, self._open)
            self._close()
        torch.cuda.current_stream().wait_event(self._done.pop(k))

    def wait_all(self):
        self._close()
        for k in sorted(self._done):
            torch.cuda.current_stream().wait_event(self._done[k])
        self._done.clear()

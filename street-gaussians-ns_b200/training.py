"""One training iteration of the scene-graph model on the fused path (BASELINE.json configs 4 and 5).

What nerfstudio's ``Trainer.train_iteration`` plus the model's training callbacks do per step for the reference
(SURVEY.md 3.1): ``step_cb`` -> zero_grad -> ``get_outputs`` -> ``get_loss_dict`` -> backward -> optimizer step ->
``after_train`` (densification statistics, sgn_splatfacto.py:513-541) -> every ``refine_every`` steps ``refinement_after``
(:550-646).  nerfstudio's engine itself (config system, data managers, viewer, checkpoints) is out of scope; this is
the sequence of hot-path calls it makes, so that a step -- and a data-parallel step -- can be run, tested and timed.

Data parallel: every replica holds all parameters and renders its own camera; the dense gradient
arena is summed with ONE all-reduce and divided by the world size, then every replica applies the same Adam update.
Replicas see different actors (different timestamps), so the arena has the layout of ALL sub-models
(``SceneGraphConfig.full_gradient_arena``) and Adam steps the UNION of the sub-models in view on any replica -- a
parameter that no replica rendered has no gradient and is skipped, as torch.optim.Adam skips ``grad is None``.  The
union is computed on the host from the cameras of all replicas (the camera assignment is deterministic,
``dp.camera_for_rank``): no extra collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import dp
from .model import SceneGraphRasterModel, _FullArenaSink
from .optim import FusedAdam
from .scene import Camera


class TrainStep:
    def __init__(self, model: SceneGraphRasterModel, optimizer: FusedAdam, refine_every: Optional[int] = None,
                 group: Optional[dist.ProcessGroup] = None, pipeline_chunks: int = 0, refine_seed: int = 0,
                 check_replicas: bool = True, overlap: bool = False, exchange: str = "auto"):
        """``pipeline_chunks`` > 0 (data parallel only): all-reduce and Adam are pipelined over that many ranges of the
        arena (dp.allreduce_and_step) instead of running one after the other."""
        self.model, self.optimizer, self.group = model, optimizer, group
        self.pipeline_chunks = pipeline_chunks if pipeline_chunks > 0 or not overlap else 4
        self.refine_seed, self.check_replicas, self._gen = refine_seed, check_replicas, None
        # how the gradient arena is exchanged: "sym" = this library's kernel over symmetric memory (csrc/collective.cu),
        # "nccl" = dist.all_reduce, "auto" = sym on NCCL process groups when the symmetric allocation succeeds
        self.exchange_mode, self._exchange = exchange, None
        # None: every sub-model on its own ``refine_every`` (the reference registers one callback per sub-model with
        # ``update_every_num_iters = config.refine_every``); a number: one cadence for all of them
        self.refine_every = refine_every
        assert optimizer.num_segments == len(model.all_models), "build FusedAdam over model.optimizer_params()"

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def submodels_in_view(self, camera: Camera) -> List[int]:
        """Indices (all_models order) of the sub-models a render of ``camera`` lists: the background and every actor with
        a box at the camera's timestamp and at least one Gaussian (scene graph :332-352)."""
        m = self.model
        index = {n: i for i, n in enumerate(m.all_models._modules)}
        seen = [0]
        for pose in m.poses_at(camera.time):
            name = m.get_object_model_name(pose.track_id)
            if m.all_models[name].num_points > 0:
                seen.append(index[name])
        return seen

    def __call__(self, step: int, camera: Camera, batch: Dict[str, torch.Tensor],
                 all_cameras: Optional[Sequence[Camera]] = None) -> Dict[str, torch.Tensor]:
        """``all_cameras``: this step's cameras of ALL replicas (rank order), needed when the world size is > 1."""
        m, opt = self.model, self.optimizer
        world = self.world_size()
        full = isinstance(m._grad_sink, _FullArenaSink)
        if world > 1:
            assert full, "data parallel needs SceneGraphConfig(full_gradient_arena=True): replicas see different actors"
            assert all_cameras is not None and len(all_cameras) == world
        if world > 1:
            self._ensure_exchange()
        m.step = step                                     # step_cb (sgn_splatfacto.py:754-755)
        for p in m.parameters():                          # Optimizers.zero_grad_all()
            p.grad = None
        out = m.get_outputs(camera)
        if world > 1 and self._exchange is not None:  # which background rows this replica sees (rows nobody sees are not exchanged)
            h = m._holder
            n_bg = self._layout[1]
            if h is not None and h.radii is not None and h.radii.shape[0] >= n_bg and m.visible_model_names[:1] == ["background"]:
                self._exchange.publish_visible(h.radii, rows=n_bg)
            else:
                self._exchange.flags.zero_()
                self._exchange._union_fresh = False
                self._exchange._union_rows = n_bg
        losses = m.get_loss_dict(out, batch)
        total = sum(losses.values())
        rendered = isinstance(total, torch.Tensor) and total.requires_grad
        if rendered:
            total.backward()
            arena = m._holder.grad_arena
        else:
            # nothing in view on this replica (the reference's early-out, sgn_splatfacto.py:878-886): no gradient here.
            # A single replica skips backward / optimizer / after_train only: the refinement callbacks still run
            # (nerfstudio fires them on the step count, whatever was rendered)
            if world == 1:
                self._maybe_refine(step)
                return losses
            arena = m.zero_gradient_arena()
        if world > 1:
            present = sorted(set(i for cam in all_cameras for i in self.submodels_in_view(cam)))
        else:
            present = m.present_submodels()
        everything = len(present) == opt.num_segments and (full or present == list(range(opt.num_segments)))
        if world > 1 and self._exchange is not None:
            self._exchange_and_step(arena, None if everything else present, rendered)
        elif world > 1 and self.pipeline_chunks > 0:
            dp.allreduce_and_step(arena, opt, None if everything else present, self.pipeline_chunks, self.group)
        else:
            if world > 1:
                dp.allreduce_gradients(arena, average=True, group=self.group)
            opt.step(arena, present=None if everything else present, full_layout=full)
        if rendered:
            m.after_train(step)                           # AFTER_TRAIN_ITERATION callbacks, in the reference's order
        self._maybe_refine(step)
        return losses

    def _ensure_exchange(self) -> None:
        """(Re)allocates the symmetric gradient arena when the model's layout changed (first step, after a refinement).
        Collective: every replica reaches it at the same step with the same sizes."""
        if self.exchange_mode == "nccl" or dist.get_backend(self.group) != "nccl":
            return
        m = self.model
        sink = m._grad_sink
        sink.bind_model(m.optimizer_params(), [])
        total = sink.total_elems()
        n_bg = m.all_models["background"].num_points
        ex = self._exchange
        layout = (total, n_bg, tuple(int(x) for x in sink.sizes[:6]))
        if ex is not None and getattr(self, "_layout", None) == layout and sink.arena is not None and sink.arena.data_ptr() == ex.arena.data_ptr():
            return
        try:
            sink.exchange_plan = None
            if ex is None or total > ex.numel or n_bg > ex.flag_rows:
                # a symmetric allocation is a collective (allocation + handle exchange + barrier): allocate with headroom so
                # that refinements -- which grow the model by a few per cent at a time -- re-use it; the faster path
                # (multimem / peer) is timed once, later allocations reuse the decision
                self._exchange = None
                ex = dp.SymmetricExchange(int(total * 1.25), m.device, self.group, flag_rows=int(n_bg * 1.25) + 128,
                                          use_multicast=getattr(self, "_use_multicast", "auto"))
                self._use_multicast = bool(ex.multicast_ptr)
            ex.arena[:total].zero_()
            sink.set_arena(ex.arena[:total])
            self._exchange, self._layout = ex, layout
            # the background's six tensors are the first six slices of the full layout: K - 1 row ranges + "everything else"
            K = max(2, self.pipeline_chunks or 4)
            widths = [3, 3, 4, 3 * int(m.all_models["background"].gauss_params["features_dc"].shape[1]),
                      3 * int(m.all_models["background"].gauss_params["features_rest"].shape[1]), 1]
            offs = [int(sink.offsets[k]) for k in range(6)]
            bg_end = int(sink.offsets[5] + sink.sizes[5])
            self._ranges = dp.plan_ranges([n_bg], [offs], [widths], K - 1)  # row ranges of the background, with row descriptions
            self._tail = (bg_end, total - bg_end)
            sink.exchange_plan = self._plan
        except Exception as e:  # no peer access / no symmetric-memory support on this box: the NCCL path is the fallback
            if self.exchange_mode == "sym":
                raise
            self._exchange = None
            self.exchange_mode = "nccl"
            self.exchange_error = f"{type(e).__name__}: {e}"[:300]

    def _plan(self, table):
        """Called by the render's backward (raster._SceneGraphRasterize.backward) with the frame's segment table: the chunk
        ranges project_bwd runs in and the callback that starts range k's exchange as soon as its launch is enqueued."""
        ex = self._exchange
        bg_chunks = self._ranges[-1][1]
        ranges = [(a, b) for a, b, _ in self._ranges]
        slices = [sl for _, _, sl in self._ranges]
        if table.num_chunks > bg_chunks or self._tail[1] > 0:  # the actors' rows of this frame, and every sub-model behind the background
            ranges.append((bg_chunks, table.num_chunks))
            slices.append([self._tail] if self._tail[1] > 0 else [])
        self._planned = slices

        def after_range(k):
            ex.after_range(k, slices[k], average=True, skip_unseen=True)
        return ranges, after_range

    def _exchange_and_step(self, arena: torch.Tensor, present, rendered: bool) -> None:
        """Mean of the arena over the replicas with sgn_allreduce_sym -- range by range behind project_bwd when this replica
        rendered (the exchanges were started by the backward), all ranges here when it did not -- and the fused Adam of range
        k as soon as range k has been exchanged (while range k+1 is on the wire)."""
        ex, opt = self._exchange, self.optimizer
        assert arena.data_ptr() == ex.arena.data_ptr(), "the gradient arena is not the symmetric allocation"
        opt.step_count += 1
        tab = opt.step_table(present, full_layout=True)
        if rendered:
            slices = self._planned
        else:  # nothing in view here: an all-zero arena, exchanged in the same ranges as on the replicas that rendered
            slices = [sl for _, _, sl in self._ranges] + ([[self._tail]] if self._tail[1] > 0 else [])
            for k, sl in enumerate(slices):
                ex.after_range(k, sl, average=True, skip_unseen=True)
        for k, sl in enumerate(slices):
            ex.wait_range(k)
            opt.launch(opt.rows_in_slices(tab, sl), arena)

    def _refine_generator(self, step: int) -> torch.Generator:
        """Split samples must be identical on every replica whatever else consumed the global CUDA generator (a sky
        map, augmentation, a different number of randn calls per rank): a dedicated generator reseeded from
        (base seed, step) before each refinement."""
        if self._gen is None:
            self._gen = torch.Generator(device=self.model.device)
        self._gen.manual_seed((self.refine_seed * 1_000_003 + step) & 0x7FFFFFFFFFFFFFFF)
        return self._gen

    def _maybe_refine(self, step: int) -> None:
        m = self.model
        if self.refine_every is not None:
            due = self.refine_every > 0 and step % self.refine_every == 0
        else:
            due = any(st.refine_every > 0 and step % st.refine_every == 0 for st in (m.config.refine, m.config.object_refine))
        if not due:
            return
        m.refinement_after(self.optimizer, step, generator=self._refine_generator(step), due_only=self.refine_every is None)
        if self.world_size() > 1 and self.check_replicas:
            # replicas must have taken identical decisions: same row count in every sub-model
            rows = torch.tensor([sub.num_points for sub in m.all_models.values()], device=m.device, dtype=torch.int64)
            lo, hi = rows.clone(), rows.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            assert torch.equal(lo, hi), "replicas diverged in a refinement (row counts differ)"


def warm_up_refinement(model: SceneGraphRasterModel, rows: Sequence[int] = (65536, 10000, 10000), step: Optional[int] = None) -> Dict[str, object]:
    """Runs ONE refinement of a throwaway model (same configuration and tensor widths as ``model``, random rows and
    statistics that straddle every threshold, its own generator) and drops it.  Nothing of ``model`` is touched and no
    collective is issued.

    Why: CUDA loads a kernel's module at its first launch.  The first refinement of a process is the first launch of ~20
    kernels nothing else on the training step uses (the decide / apply kernels, scans, reductions, the normal sampler,
    bitwise and comparison kernels, ...), which cost tens of milliseconds of host time in the middle of training
    (profiles/r02p_refine_profile_*: 60 ms for a refinement whose kernels run a few ms).  Calling this once after the
    model is built moves that one-time cost out of the loop, like the warm-up steps of a benchmark."""
    from dataclasses import replace

    from .scene import GaussianSet
    dev = model.device
    subs = list(model.all_models.values())
    widths = [subs[0]] + [s for s in subs[1:]][:len(rows) - 1]
    g = torch.Generator(device=dev).manual_seed(12345)

    def rand(*shape):
        return torch.randn(shape, device=dev, generator=g)

    sets = []
    for n, like in zip(rows, widths):
        F, R = int(like.gauss_params["features_dc"].shape[1]), int(like.gauss_params["features_rest"].shape[1])
        sets.append(GaussianSet(rand(n, 3) * 5, rand(n, 3) * 1.5 - 4.0, rand(n, 4), rand(n, F, 3), rand(n, R, 3), rand(n, 1) * 2.5 - 1.0))
    st = model.config.refine
    step = step if step is not None else st.warmup_length + st.refine_every * max(1, -(-(model.config.num_train_data + st.refine_every + 1) // st.refine_every))
    cfg = replace(model.config, full_gradient_arena=False, refine_record=True)
    tmp = SceneGraphRasterModel(sets[0], {str(i): s for i, s in enumerate(sets[1:])}, cfg).to(dev)
    tmp.train()
    tmp.step = step
    for sub in tmp.all_models.values():
        n, d = sub.num_points, sub.__dict__
        vis = torch.randint(1, 9, (n,), device=dev, generator=g).float()
        d["vis_counts"] = vis
        d["xys_grad_norm"] = torch.rand(n, device=dev, generator=g) * vis * 2.5e-6
        d["max_2Dsize"] = torch.rand(n, device=dev, generator=g) * 0.2
        d["last_size"] = (240, 320)
    opt = FusedAdam(tmp.optimizer_params())
    tmp.refinement_after(opt, step, generator=g, sync_stats=False)
    rowsum = torch.tensor([s.num_points for s in tmp.all_models.values()], device=dev, dtype=torch.int64)
    torch.equal(rowsum, rowsum.clone())  # the replica check of TrainStep._maybe_refine
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    return {"step": step, "rows_before": [int(n) for n in rows[:len(sets)]], "rows_after": rowsum.tolist(),
            "records": [dict(s.refine_record_dict) for s in tmp.all_models.values()]}

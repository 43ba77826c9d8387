"""Level-2 drop-in surface: the nerfstudio ``Model`` methods of the reference's scene-graph model
that sit on the hot path, re-hosted on the fused rasterizer.

Mirrors ``SplatfactoSceneGraphModel`` (street_gaussians_ns/sgn_splatfacto_scene_graph.py:41-401) and the
parts of ``SplatfactoModel`` it inherits (street_gaussians_ns/sgn_splatfacto.py:793-1001, :1042-1094):

  * ``all_models`` ModuleDict: "background", "object_<id>" ... each holding the six ``gauss_params``
    (names/shapes unchanged, so reference checkpoints load: ``all_models.<name>.gauss_params.<p>``);
  * ``get_outputs(camera)`` -> {"rgb","accumulation","depth","sky","object_acc","background_acc"
    (+ "background_rgb","object_rgb" in eval)} with the reference's early-outs;
  * side-effect attributes ``xys`` (with ``.grad`` = pixel-space mean gradient after backward),
    ``depths, radii, conics, num_tiles_hit, last_size`` on the model and split per sub-model
    (``SplatfactoModel.after_train`` reads ``self.xys.grad`` and ``self.radii``, :513-541);
  * ``get_loss_dict`` (L1 + SSIM + sky accumulation + object-accumulation entropy).

nerfstudio is not a dependency here: ``camera`` is any object with the fields of
``street_gaussians_ns_b200.scene.Camera`` (SURVEY.md 8b lists what the path reads from ``Cameras``).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import raster, refine
from .refine import RefineSettings
from .scene import PARAM_NAMES, CLS_BACKGROUND, CLS_OBJECT, Camera, Frame, GaussianSet, Segment, fourier_time, idft_basis


@dataclass
class ActorPose:
    """What the path needs from a reference ``Box`` (data/utils/dynamic_annotation.py): trackId,
    center, rot, frame; plus the actor's frame list for the Fourier time (scene graph :239-245)."""

    track_id: str
    rot: np.ndarray
    center: np.ndarray
    frame: int
    frame_list: Sequence[int]


@dataclass
class SceneGraphConfig:
    sh_degree: int = 3
    sh_degree_interval: int = 1000
    block_width: int = 16
    fourier_features_dim: int = 5
    fourier_features_scale: float = 1.0
    use_sky_sphere: bool = True
    ssim_lambda: float = 0.2
    sky_acc_loss_mult: float = 1.0
    object_acc_entropy_loss_mult: float = 0.001
    alpha_clamp_fwd: float = 0.999
    alpha_clamp_bwd: float = 0.99
    render_background_acc: bool = True
    fused_loss: bool = True  # L1 / sky / entropy terms through the fused loss epilogue (loss.py) instead of torch ops
    # refinement (sgn_splatfacto.py:550-646): the sub-model configs of sgn_config.py:46-66 (background / object template)
    refine: RefineSettings = field(default_factory=lambda: RefineSettings(cull_alpha_thresh=0.02))
    object_refine: RefineSettings = field(default_factory=lambda: RefineSettings(cull_alpha_thresh=0.005))
    num_train_data: int = 0      # Model.num_train_data: refinement waits until every image was seen after an opacity reset (:563-566)
    refine_record: bool = False  # fill sub.refine_record_dict (the reference's logging counters; one extra read-back per sub-model)
    # data parallel: gradients are delivered in an arena that has the layout of ALL sub-models (the
    # optimizer's), so that replicas which see different actors can sum their arenas with one all-reduce
    full_gradient_arena: bool = False
    # no host read-back of the intersection count inside get_outputs (raster.RenderSettings.async_binning); None -> SGN_ASYNC_BIN
    async_binning: Optional[bool] = None

    # ``stop_split_at`` of the BACKGROUND sub-model: what the reference's entropy gate reads
    # (``config.background_model.stop_split_at``, scene graph :386).  One source: ``refine.stop_split_at``.
    @property
    def stop_split_at(self) -> int:
        return self.refine.stop_split_at

    @stop_split_at.setter
    def stop_split_at(self, v: int) -> None:
        self.refine.stop_split_at = int(v)


_MISSING = object()


class _FrameSlice:
    """A per-frame side-effect attribute of a sub-model (``xys``, ``depths``, ``radii``, ``conics``, ``num_tiles_hit``;
    sgn_splatfacto.py:513-541 reads them): this sub-model's rows of the frame's arrays, sliced on first access.  Publishing
    them eagerly costs 5 slices x 33 sub-models of host time per frame, and nothing on the hot path reads them (the
    densification statistics work on the frame's arrays directly)."""

    def __init__(self, name: str):
        self.name = name

    def __get__(self, obj, cls=None):
        if obj is None:
            return self
        d = obj.__dict__
        cache = d.get("_fs_cache")
        if cache is not None:
            v = cache.get(self.name, _MISSING)
            if v is not _MISSING:
                return v
        src = d.get("_fs_src")
        if src is None:
            return None
        holder, sl = src
        t = getattr(holder, self.name)[sl]
        if self.name == "xys" and holder.v_records is not None:  # after backward: the reference reads ``self.xys.grad``
            t.grad = holder.v_records[:, 0:2][sl]
        if cache is None:
            cache = d["_fs_cache"] = {}
        cache[self.name] = t
        return t

    def __set__(self, obj, value):
        cache = obj.__dict__.get("_fs_cache")
        if cache is None:
            cache = obj.__dict__["_fs_cache"] = {}
        cache[self.name] = value


class GaussianSubModel(torch.nn.Module):
    """One entry of ``all_models``: the ``gauss_params`` ParameterDict (sgn_splatfacto.py:291-300)."""

    xys = _FrameSlice("xys")
    depths = _FrameSlice("depths")
    radii = _FrameSlice("radii")
    conics = _FrameSlice("conics")
    num_tiles_hit = _FrameSlice("num_tiles_hit")

    def __init__(self, params: GaussianSet):
        super().__init__()
        self.gauss_params = torch.nn.ParameterDict(
            {k: torch.nn.Parameter(getattr(params, k).detach().clone()) for k in
             ("means", "scales", "quats", "features_dc", "features_rest", "opacities")})
        self.xys = self.depths = self.radii = self.conics = self.num_tiles_hit = None
        self.last_size = None
        # densification statistics (sgn_splatfacto.py:513-541), created by the first after_train
        self.xys_grad_norm = self.vis_counts = self.max_2Dsize = None
        self.refine_record_dict: dict = {}

    @property
    def num_points(self) -> int:
        return int(self.gauss_params["means"].shape[0])

    def as_set(self) -> GaussianSet:
        g = self.gauss_params
        return GaussianSet(g["means"], g["scales"], g["quats"], g["features_dc"], g["features_rest"], g["opacities"])

    def load_state_dict(self, state_dict, **kwargs):  # type: ignore[override]
        """Checkpoints hold whatever number of Gaussians refinement had reached: resize the parameters to the
        checkpoint's row count before loading, and accept the pre-ParameterDict key names
        (sgn_splatfacto.py:425-438)."""
        state_dict = dict(state_dict)
        if "means" in state_dict:
            for k in PARAM_NAMES:
                state_dict[f"gauss_params.{k}"] = state_dict.pop(k)
        if "gauss_params.means" in state_dict:
            rows = state_dict["gauss_params.means"].shape[0]
            for k in PARAM_NAMES:
                old = self.gauss_params[k]
                want = state_dict.get(f"gauss_params.{k}")
                shape = (rows,) + (tuple(want.shape[1:]) if want is not None else tuple(old.shape[1:]))
                if tuple(old.shape) != shape:
                    self.gauss_params[k] = torch.nn.Parameter(torch.zeros(shape, device=old.device, dtype=old.dtype))
            d = self.__dict__
            d["xys_grad_norm"] = d["vis_counts"] = d["max_2Dsize"] = None  # statistics of the old rows are meaningless
        return super().load_state_dict(state_dict, **kwargs)


class _GradSink:
    """Delivers the rasterizer's flat gradient arena to the model's ~200 parameter tensors.

    torch.autograd with one leaf per parameter tensor costs ~1.5 ms of host time per step here (198
    AccumulateGrad nodes + 198 views), more than half of the GPU time of the whole step.  Instead the
    render has ONE differentiable input (an anchor), and after backward the sink points every
    ``param.grad`` at its slice of a persistent arena that project_bwd overwrites in place.
    Semantics match autograd's: a parameter whose ``.grad`` is None (after ``zero_grad()``) gets the new
    gradient; if any gradient is still set (no zero_grad between backward calls) the new one is ADDED."""

    def __init__(self):
        self.key = None
        self.arena = None
        self.views: List[torch.Tensor] = []
        self.params: List[torch.Tensor] = []

    def bind(self, params: List[torch.Tensor]):
        key = tuple(map(id, params))
        if key != self.key:
            self.key, self.params, self.arena, self.views = key, list(params), None, []

    allocator = None  # optional callable(total, device) -> float32 tensor: where the persistent arena lives (data parallel: a
    #                   symmetric allocation the exchange kernel can address on every replica)

    def target(self, static: dict, device):
        sizes, _, _ = raster.arena_layout(static)
        total = sum(sizes)
        if self.arena is None or self.arena.numel() != total or self.arena.device != device:
            self.arena = self.allocator(total, device) if self.allocator is not None else torch.empty(total, device=device, dtype=torch.float32)
            assert self.arena.numel() == total and self.arena.dtype == torch.float32
            self.views = raster.arena_views(self.arena, static)
        for p in self.params:
            if p.grad is not None:
                return None  # accumulate: render into a temporary arena, add in publish()
        return self.arena

    def grad_offsets(self, static: dict):
        """Where project_bwd puts the frame's slices inside ``target()``'s arena: None = back to back (the frame's layout)."""
        return None

    def publish(self, arena: torch.Tensor, static: dict):
        if arena is self.arena:
            for p, v in zip(self.params, self.views):
                p.grad = v
            return
        for p, v, pv in zip(self.params, raster.arena_views(arena, static), self.views):
            if p.grad is None:
                pv.copy_(v)
                p.grad = pv
            else:
                p.grad.add_(v)


class _FullArenaSink(_GradSink):
    """Data-parallel form of the sink: the persistent arena has the layout of ALL sub-models (the optimizer's layout,
    ``FusedAdam(model.optimizer_params())``), whatever subset is in view.  Replicas render different cameras at
    different timestamps and therefore see different sets of actors; only with a common layout can ONE all-reduce sum
    their gradients.  The slices of sub-models that are not in this replica's frame are zeroed (their sum over the
    replicas is then the other replicas' gradient), the frame's slices are written in place by project_bwd through
    per-tensor offsets."""

    def __init__(self):
        super().__init__()
        self.model_key = None
        self.sizes = self.offsets = None   # per tensor of the whole model (floats, padded to 4)
        self.present: List[int] = []
        self.all_views: List[torch.Tensor] = []
        self.all_shapes: List[tuple] = []

    def bind_model(self, all_params: List[List[torch.Tensor]], present: List[int]):
        key = tuple((id(t), t.shape[0]) for ps in all_params for t in ps)  # ids can be recycled after a refinement
        if key != self.model_key:
            flat = [t for ps in all_params for t in ps]
            self.model_key, self.arena, self.all_views = key, None, []
            self.all_shapes = [tuple(t.shape) for t in flat]
            self.sizes = np.array([(t.numel() + 3) // 4 * 4 for t in flat], np.int64)
            self.offsets = np.concatenate([[0], np.cumsum(self.sizes)[:-1]]).astype(np.int64)
        self.present = list(present)
        self.params = [t for i in self.present for t in all_params[i]]
        self.key = tuple(map(id, self.params))

    def _tensor_ids(self):
        return [6 * i + k for i in self.present for k in range(6)]

    def grad_offsets(self, static: dict):
        return self.offsets[self._tensor_ids()]

    def total_elems(self) -> int:
        return int(self.sizes.sum())

    def set_arena(self, arena: torch.Tensor) -> None:
        """Install the arena the gradients are delivered in (data parallel: a view of a symmetric allocation, so that the
        exchange kernel of csrc/collective.cu can address every replica's copy)."""
        assert arena.dtype == torch.float32 and arena.is_contiguous() and arena.numel() == self.total_elems()
        self.arena = arena
        chunks = arena.split_with_sizes([int(x) for x in self.sizes])
        self.all_views = [c[:int(np.prod(shp))].view(shp) for c, shp in zip(chunks, self.all_shapes)]

    def target(self, static: dict, device):
        total = int(self.sizes.sum())
        if self.arena is None or self.arena.numel() != total or self.arena.device != device:
            self.set_arena(torch.zeros(total, device=device, dtype=torch.float32))
        self.views = [self.all_views[t] for t in self._tensor_ids()]
        for p in self.params:
            if p.grad is not None:
                return None  # accumulate: render into a temporary arena (frame layout), add in publish()
        # zero the sub-models that are not in this frame (runs of absent sub-models are contiguous in the arena);
        # after an all-reduce they hold the other replicas' gradients
        nsub = len(self.sizes) // 6
        here = set(self.present)
        i = 0
        while i < nsub:
            if i in here:
                i += 1
                continue
            j = i
            while j < nsub and j not in here:
                j += 1
            lo = int(self.offsets[6 * i])
            hi = int(self.offsets[6 * (j - 1) + 5] + self.sizes[6 * (j - 1) + 5])
            self.arena[lo:hi].zero_()
            i = j
        return self.arena


_MODEL_SERIAL = itertools.count()


class SceneGraphRasterModel(torch.nn.Module):
    def __init__(self, background: GaussianSet, actors: Dict[str, GaussianSet], config: Optional[SceneGraphConfig] = None,
                 poses_at: Optional[Callable[[float], List[ActorPose]]] = None,
                 sky: Optional[Callable[[Camera, bool], torch.Tensor]] = None):
        super().__init__()
        self.config = config or SceneGraphConfig()
        self.all_models = torch.nn.ModuleDict()
        self.all_models["background"] = GaussianSubModel(background)
        for obj_id, ps in actors.items():
            self.all_models[self.get_object_model_name(obj_id)] = GaussianSubModel(ps)
        self.poses_at = poses_at or (lambda t: [])
        self.env_map = sky  # stays on nvdiffrast in the reference (EnvLight, sgn_splatfacto.py:109-150)
        self.step = 0
        self.visible_model_names: List[str] = ["background"]
        self.xys = self.depths = self.radii = self.conics = self.num_tiles_hit = None
        self.last_size = None
        self._holder = None
        self._frame_cache: dict = {}
        self.__dict__["_serial"] = next(_MODEL_SERIAL)  # names this model in raster's static segment-table cache (ids are recycled)
        self._grad_sink = _FullArenaSink() if self.config.full_gradient_arena else _GradSink()
        self._anchor = None

    @staticmethod
    def get_object_model_name(object_id) -> str:
        return f"object_{object_id}"

    def load_state_dict(self, state_dict, **kwargs):  # type: ignore[override]
        """``SplatfactoSceneGraphModel.load_state_dict`` (scene graph :393-401): every sub-model takes the keys under
        ``all_models.<name>.`` (and resizes itself to the checkpoint's row count), the rest loads non-strictly."""
        state_dict = dict(state_dict)
        for name, sub in self.all_models.items():
            prefix = f"all_models.{name}."
            own = {k[len(prefix):]: state_dict.pop(k) for k in list(state_dict) if k.startswith(prefix)}
            if own:
                sub.load_state_dict(own, **kwargs)
        self.invalidate_frames()
        return torch.nn.Module.load_state_dict(self, state_dict, strict=False)

    @property
    def device(self):
        return self.all_models["background"].gauss_params["means"].device

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): parameter storage moves, staged pointers are stale
        out = super()._apply(fn, *args, **kwargs)
        if "_frame_cache" in self.__dict__:
            self.invalidate_frames()
        return out

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _pose_digest(poses) -> tuple:
        """Content key of a timestamp's boxes: an integration may move boxes at a fixed timestamp (the reference's
        ``bbox_optimizer.apply_to_bbox`` rewrites them every training step) or hand out fresh objects whose ``id`` CPython
        recycles -- identity says nothing, the bytes do."""
        return tuple((p.track_id, p.frame, len(p.frame_list), p.frame_list[0], p.frame_list[-1],
                      np.asarray(p.rot, np.float64).tobytes(), np.asarray(p.center, np.float64).tobytes()) for p in poses)

    def _frame(self, camera: Camera) -> Frame:
        """The Frame of ``camera``.  The segment rows of a timestamp (poses, IDFT bases, parameter pointers, row offsets) are
        static data while the parameter tensors stay the same objects and the boxes keep their values: they are kept per
        timestamp -- device-resident when ``prepare_frames`` built them up front -- and validated by content."""
        poses = self.poses_at(camera.time)
        digest = self._pose_digest(poses)
        hit = self._frame_cache.get(camera.time)
        if hit is not None and hit["digest"] == digest and hit["ptr0"] == self.all_models._modules["background"].gauss_params._parameters["means"].data_ptr():
            self.visible_model_names = hit["names"]
            frame = Frame(camera, hit["segments"])
            frame._prebuilt, frame._table_slot = hit["table"], hit
            frame._static_key = ("model", self.__dict__["_serial"], self.__dict__.get("_param_epoch", 0), tuple(hit["names"]), str(self.device))
            return frame
        frame = self._build_frame(camera, poses)
        frame._table_slot = self._remember_frame(camera.time, digest, frame, None)
        return frame

    def invalidate_frames(self) -> None:
        """Drop the per-timestamp segment rows (call after replacing parameter tensors by hand; ``refinement_after`` and
        ``load_state_dict`` do it themselves)."""
        self._frame_cache.clear()
        self.__dict__.pop("_frame_table_blob", None)
        self.__dict__["_param_epoch"] = self.__dict__.get("_param_epoch", 0) + 1
        self.__dict__["_sets"] = {}

    def _remember_frame(self, time, digest, frame: Frame, table):
        if len(self._frame_cache) > 4096:
            self._frame_cache.clear()
        slot = self._frame_cache[time] = dict(
            digest=digest, segments=frame.segments, names=list(self.visible_model_names), table=table,
            ptr0=self.all_models._modules["background"].gauss_params._parameters["means"].data_ptr())
        return slot

    def prepare_frames(self, times: Sequence[float]) -> int:
        """the segment table of EVERY (timestamp, actor) -- object->world rotation / translation /
        quaternion (``object2world_gs``, scene graph :404-417), IDFT basis (:420-433), Fourier time (:239-245), parameter
        pointers and row offsets -- built once from the annotation table (``poses_at``; for the reference's
        ``InterpolatedAnnotation`` that includes slerp-interpolated boxes with ``frame = -1``,
        dynamic_annotation.py:157-171,252-286) and uploaded with ONE copy.  ``get_outputs`` then indexes it by timestamp: no
        per-frame host build, no per-frame H2D.  Rebuilt by the caller after a refinement replaced parameter tensors
        (the cache is dropped there).  Returns the bytes resident on the device."""
        dev = self.device
        built = []
        for t in times:
            poses = self.poses_at(t)
            frame = self._build_frame(None, poses)
            params = [list(seg.params.tensors()) for seg in frame.segments]
            tab = raster.SegmentTable(frame, params, dev, upload=False)
            built.append((t, self._pose_digest(poses), frame, tab))
        if not built:
            return 0
        blob = np.concatenate([tab.host.view(np.uint8).reshape(-1) for _, _, _, tab in built])
        resident = torch.from_numpy(blob).to(dev)
        off = 0
        for t, digest, frame, tab in built:
            nbytes = tab.host.nbytes
            tab.dev = resident[off:off + nbytes]
            off += nbytes
            self.visible_model_names = [seg.name for seg in frame.segments]
            self._remember_frame(t, digest, frame, tab)
        self.__dict__["_frame_table_blob"] = resident
        return int(resident.numel())

    def _set_of(self, name: str) -> GaussianSet:
        sets = self.__dict__.setdefault("_sets", {})
        g = sets.get(name)
        if g is None:
            g = sets[name] = self.all_models._modules[name].as_set()
        return g

    def _build_frame(self, camera: Camera, poses) -> Frame:
        segs = [Segment(self._set_of("background"), CLS_BACKGROUND, name="background")]
        names = ["background"]
        for pose in poses:
            name = self.get_object_model_name(pose.track_id)
            assert name not in names
            ps = self._set_of(name)
            if ps.means.shape[0] == 0:  # "prevent empty object" (scene graph :337-338)
                continue
            basis = None
            F = ps.features_dc.shape[1]
            if F > 1:
                t = fourier_time(pose.frame, pose.frame_list, self.config.fourier_features_scale)
                basis = idft_basis(t, F)
            segs.append(Segment(ps, CLS_OBJECT, rot=pose.rot, center=pose.center, idft=basis, name=name))
            names.append(name)
        self.visible_model_names = names
        frame = Frame(camera, segs)
        frame._static_key = ("model", self.__dict__["_serial"], self.__dict__.get("_param_epoch", 0), tuple(names), str(self.device))
        return frame

    def _settings(self, class_streams: bool) -> raster.RenderSettings:
        c = self.config
        n = min(self.step // c.sh_degree_interval, c.sh_degree) if self.training else c.sh_degree
        return raster.RenderSettings(sh_degree=c.sh_degree, sh_degree_to_use=n, block_width=c.block_width,
                                     alpha_clamp_fwd=c.alpha_clamp_fwd, alpha_clamp_bwd=c.alpha_clamp_bwd,
                                     class_streams=class_streams, training=self.training, async_binning=c.async_binning)

    def get_outputs(self, camera: Camera) -> Dict[str, torch.Tensor]:
        """``SplatfactoSceneGraphModel.get_outputs`` (scene graph :305-374)."""
        assert camera.time is not None
        frame = self._frame(camera)
        H, W = camera.height, camera.width
        self.last_size = (H, W)
        sky = self.env_map(camera, self.training) if (self.config.use_sky_sphere and self.env_map is not None) else None
        sink = anchor = None
        if torch.is_grad_enabled():
            flat = [t for seg in frame.segments for t in seg.params.tensors()]
            if all(t.requires_grad and t.is_leaf for t in flat):  # the normal case: the model's own nn.Parameters
                sink = self._grad_sink
                if isinstance(sink, _FullArenaSink):
                    sink.bind_model(self.optimizer_params(), self.present_submodels())
                else:
                    sink.bind(flat)
                if self._anchor is None or self._anchor.device != flat[0].device:
                    self._anchor = torch.zeros(1, device=flat[0].device, requires_grad=True)
                anchor = self._anchor
        out, holder = raster.render_frame(frame, self._settings(class_streams=True), sky=sky, grad_sink=sink, anchor=anchor)
        self._holder = holder
        self._publish_side_effects(frame, holder)
        if isinstance(holder.M, raster.LazyCount):
            # rendered without the count's read-back (RenderSettings.async_binning): the reference's early-out for "nothing in
            # view" (depth 0 instead of the 10 of an empty pixel, sgn_splatfacto.py:878-886) is applied on the device
            out["depth"] = out["depth"] * (holder.M.device_count > 0).to(out["depth"].dtype)
        elif holder.M == 0:
            # reference early-out when nothing is visible (sgn_splatfacto.py:878-886): background colour
            # (zeros), zero accumulation and ZERO depth
            dev = self.device
            res = {"rgb": torch.zeros(H, W, 3, device=dev), "accumulation": torch.zeros(H, W, 1, device=dev),
                   "depth": torch.zeros(H, W, 1, device=dev)}
            if sky is not None:
                res["sky"] = sky
            res["object_acc"] = torch.zeros(H, W, 1, device=dev)
            res["background_acc"] = torch.zeros(H, W, 1, device=dev)
            return res
        if not self.training:
            # eval-only extra renders (scene graph :367-372): per-class rgb
            with torch.no_grad():
                out["background_rgb"] = self._class_rgb(frame, CLS_BACKGROUND, sky)
                out["object_rgb"] = self._class_rgb(frame, CLS_OBJECT, None)
                if not any(s.cls == CLS_OBJECT for s in frame.segments):
                    # without actors the reference's objects-only render returns {'rgb': zeros[H,W,1], 'depth': zeros[H,W,1]}
                    # (scene graph :264-267), which get_outputs publishes as object_rgb AND object_depth (:371-372)
                    out["object_depth"] = torch.zeros(H, W, 1, device=self.device)
        return out

    def _class_rgb(self, frame: Frame, cls: int, sky):
        segs = [s for s in frame.segments if s.cls == cls]
        H, W = frame.camera.height, frame.camera.width
        if not segs:
            return torch.zeros(H, W, 1, device=self.device) if sky is None else sky
        sub, _ = raster.render_frame(Frame(frame.camera, segs), self._settings(class_streams=False), sky=sky)
        return sub["rgb"]

    def _refine_settings_of(self, sub) -> RefineSettings:
        return self.config.refine if sub is self.all_models._modules["background"] else self.config.object_refine

    def _publish_side_effects(self, frame: Frame, holder) -> None:
        # plain tensors, not parameters/buffers: write the instance dicts directly (nn.Module.__setattr__ costs
        # ~5 us per attribute, x6 attributes x33 sub-models per frame)
        d = self.__dict__
        d["xys"], d["depths"], d["radii"] = holder.xys, holder.depths, holder.radii
        d["conics"], d["num_tiles_hit"] = holder.conics, holder.num_tiles_hit
        mods = self.all_models._modules
        visible = set(self.visible_model_names)
        for name, sub in mods.items():
            if name not in visible:
                sd = sub.__dict__
                sd["_fs_src"], sd["_fs_cache"] = None, None
        row = 0
        slices = []
        for seg in frame.segments:
            sub = mods[seg.name]
            n = seg.params.means.shape[0]
            sl = slice(row, row + n)
            sd = sub.__dict__
            sd["_fs_src"], sd["_fs_cache"], sd["last_size"] = (holder, sl), None, self.last_size  # sliced on first access
            slices.append((sub, sl))
            row += n
        d["_slices"] = slices
        holder.post_backward = self._split_xys_grad(slices)

    def after_train(self, step: int) -> None:
        """The ``after_train`` callbacks of all visible sub-models (sgn_splatfacto.py:513-541; the scene graph
        registers one per sub-model, :127-137) as one launch of ``sgn_densify_stats`` over the frame's rows:
        running ||xys.grad|| sums, visibility counts and the max screen-space radius ratio, per sub-model."""
        assert step == self.step
        h = self._holder
        if h is None or h.v_records is None:
            return
        import ctypes as C
        from . import _lib
        L = _lib.load()
        # every sub-model's callback reads ITS OWN config.stop_split_at (sgn_splatfacto.py:516-518)
        slices = [(sub, sl) for sub, sl in (self.__dict__.get("_slices") or [])
                  if self.step < self._refine_settings_of(sub).stop_split_at]
        if not slices:
            return
        tab = (_lib.DensifySegment * len(slices))()
        dev = h.v_records.device
        for j, (sub, sl) in enumerate(slices):
            n = sl.stop - sl.start
            first = sub.xys_grad_norm is None or sub.xys_grad_norm.shape[0] != n
            if first:
                d = sub.__dict__
                d["xys_grad_norm"], d["vis_counts"], d["max_2Dsize"] = (torch.empty(n, device=dev) for _ in range(3))
            tab[j].row0, tab[j].count, tab[j].first = sl.start, n, int(first)
            tab[j].xys_grad_norm, tab[j].vis_counts = sub.xys_grad_norm.data_ptr(), sub.vis_counts.data_ptr()
            tab[j].max_2Dsize = sub.max_2Dsize.data_ptr()
        raw = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev, non_blocking=True)
        H, W = self.last_size
        _lib.check(L.sgn_densify_stats(raster._ptr(raw), len(slices), h.v_records.shape[0], raster._ptr(h.v_records),
                                       raster._ptr(h.radii), H, W, raster._stream()), "sgn_densify_stats")

    # ------------------------------------------------------------------------------------------
    def refinement_after(self, optimizers, step: int, generator: Optional[torch.Generator] = None, sync_stats: bool = True,
                         due_only: bool = False) -> None:
        """The ``refinement_after`` callbacks of all sub-models (sgn_splatfacto.py:550-646; the scene graph registers one
        per sub-model, :127-137): split / duplicate / cull / opacity reset, with the optimizer state carried along
        (dup_in_optim / remove_from_optim, :459-511).  ``optimizers`` is a ``FusedAdam`` built over
        ``self.optimizer_params()``, or the reference's form -- an object (or dict) mapping the six group names to
        ``torch.optim.Adam`` instances whose ``param_groups[0]["params"][i]`` is sub-model i's tensor -- or None.

        Two phases so that nothing is copied twice and the host waits once: decide every sub-model (all launches first,
        then ONE read-back of all the counts), then lay out the new tensors / moment arenas and let ``sgn_refine_apply``
        write each sub-model's rows into them.

        ``due_only``: nerfstudio runs each sub-model's callback every ITS OWN ``refine_every`` steps
        (``update_every_num_iters=self.config.refine_every``, sgn_splatfacto.py:773-781); with ``due_only`` a sub-model
        whose ``refine_every`` does not divide ``step`` is left alone (callers that run on a global cadence pass False)."""
        assert step == self.step
        names = list(self.all_models._modules)
        subs = [self.all_models[n] for n in names]
        adapter = _optimizer_adapter(optimizers, len(subs))
        if sync_stats:
            self._sync_densification_stats(subs)
        # phase 1: every sub-model's decision pass is launched, then ALL totals come back with one read-back
        decided, resets = [], []
        for name, sub in zip(names, subs):
            st = self.config.refine if name == "background" else self.config.object_refine
            densify, cull_only, reset = refine.phase(st, step, self.config.num_train_data)
            if due_only and (st.refine_every <= 0 or step % st.refine_every != 0):
                decided.append(None)
                resets.append(None)
                continue
            sub.__dict__["refine_record_dict"] = {}
            if step <= st.warmup_length or sub.xys_grad_norm is None:  # :552-555
                decided.append(None)
                resets.append(None)
                continue
            entry = None
            if densify or cull_only:
                size = sub.last_size or self.last_size
                cfg = refine.make_config(st, step, size, densify)
                g = sub.gauss_params
                flags, scan = refine.decide_submodel(g["scales"].data, g["opacities"].data, sub.xys_grad_norm if densify else None,
                                                     sub.vis_counts if densify else None,
                                                     sub.max_2Dsize if cfg.use_screen_size else None, cfg)
                entry = (flags, scan, cfg)
            decided.append(entry)
            resets.append(st if reset else None)
            d = sub.__dict__
            d["xys_grad_norm"] = d["vis_counts"] = d["max_2Dsize"] = None  # :644-646
        totals = iter(refine.read_totals([e[1] for e in decided if e is not None]))
        plans, records = [], []
        for sub, entry in zip(subs, decided):  # split samples are drawn in sub-model order, as the reference's callbacks run
            plan = None
            if entry is not None:
                plan = refine.finish_plan(entry[0], entry[1], next(totals), entry[2], generator)
                if self.config.refine_record:
                    records.append((sub, plan, plan.record_counts()))
                if not plan.changed:
                    plan = None
            plans.append(plan)
        old = [[sub.gauss_params[k].data for k in PARAM_NAMES] for sub in subs]
        new = [o if p is None else [torch.empty((p.out_rows,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype) for t in o]
               for o, p in zip(old, plans)]
        if any(p is not None for p in plans):
            src_m, dst_m = adapter.relayout(new, [p is not None for p in plans])
            for i, p in enumerate(plans):
                if p is not None:
                    refine.apply_plan(p, old[i], new[i], src_m[i], dst_m[i])
                    for k, t in zip(PARAM_NAMES, new[i]):
                        subs[i].gauss_params[k] = torch.nn.Parameter(t)
            adapter.commit([[sub.gauss_params[k] for k in PARAM_NAMES] for sub in subs], [p is not None for p in plans])
            self.invalidate_frames()
        for i, st in enumerate(resets):
            if st is not None:  # opacity reset on the survivors (:629-642)
                subs[i].gauss_params["opacities"].data.clamp_(max=refine.opacity_reset_logit(st))
                adapter.zero_moments(i, 5)
        if records:  # the logged counters of all sub-models: one read-back, after everything else has been enqueued
            for (sub, plan, _), counts in zip(records, torch.stack([r[2] for r in records]).tolist()):
                sub.__dict__["refine_record_dict"] = plan.record_from(counts)

    def _sync_densification_stats(self, subs) -> None:
        """Replicas rendered different cameras: identical split / cull decisions need identical statistics.
        SUM / SUM / MAX per sub-model; a replica that has not seen a sub-model since the last refinement contributes
        zeros, and a sub-model nobody saw stays without statistics (the collectives are the same on every replica)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        dev = self.device
        has = torch.tensor([float(sub.xys_grad_norm is not None) for sub in subs], device=dev)
        dist.all_reduce(has, op=dist.ReduceOp.MAX)
        live = []
        for sub, h in zip(subs, has.tolist()):
            if not h:
                continue
            if sub.xys_grad_norm is None:
                d = sub.__dict__
                d["xys_grad_norm"], d["vis_counts"], d["max_2Dsize"] = (torch.zeros(sub.num_points, device=dev) for _ in range(3))
                d["last_size"] = sub.last_size or self.last_size
            live.append(sub)
        if not live:
            return
        # TWO collectives for all sub-models (33 x 3 small ones cost ~3 ms per refinement at eight GPUs): the SUM statistics and the
        # MAX statistic are packed into flat buffers, reduced, and copied back
        sums = torch.cat([t for sub in live for t in (sub.xys_grad_norm, sub.vis_counts)])
        maxs = torch.cat([sub.max_2Dsize for sub in live])
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
        o = m_ = 0
        for sub in live:
            n = sub.num_points
            sub.xys_grad_norm.copy_(sums[o:o + n])
            sub.vis_counts.copy_(sums[o + n:o + 2 * n])
            sub.max_2Dsize.copy_(maxs[m_:m_ + n])
            o += 2 * n
            m_ += n

    def zero_gradient_arena(self) -> torch.Tensor:
        """Data parallel: this replica rendered nothing (early-out) but the others did -- its contribution to the
        all-reduce is an all-zero arena in the common layout."""
        sink = self._grad_sink
        assert isinstance(sink, _FullArenaSink), "only with SceneGraphConfig(full_gradient_arena=True)"
        sink.bind_model(self.optimizer_params(), [])
        return sink.target(None, self.device)

    def optimizer_params(self) -> List[List[torch.Tensor]]:
        """Per sub-model (all_models order: background, then objects), the six tensors in gradient-arena order: what
        ``FusedAdam`` is built over.  ``present_submodels()`` names the ones in the last frame's gradient arena."""
        return [[sub.gauss_params[k] for k in PARAM_NAMES] for sub in self.all_models._modules.values()]

    def present_submodels(self) -> List[int]:
        index = {n: i for i, n in enumerate(self.all_models._modules)}
        return [index[n] for n in self.visible_model_names]

    @staticmethod
    def _split_xys_grad(slices):
        """After ``loss.backward()``: ``sub.xys.grad`` for every visible sub-model, as the reference's
        ``set_split_tensor_variable(..., retain_grad=True)`` provides (scene graph :153-179).  ``sub.xys`` is sliced on access
        (with its gradient, once the backward has run); only a ``sub.xys`` that was read BEFORE the backward needs it set here."""
        def hook(h):
            v_xy = None
            for sub, sl in slices:
                cache = sub.__dict__.get("_fs_cache")
                if cache:
                    t = cache.get("xys")
                    if t is not None:
                        v_xy = h.v_records[:, 0:2] if v_xy is None else v_xy
                        t.grad = v_xy[sl]
        return hook

    # ------------------------------------------------------------------------------------------
    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        """sgn_splatfacto.py:1042-1094 + scene graph :376-391.  ``batch["image"]`` may be the float image or the
        uint8 one the data loader holds (converted as the reference does: ``.float() / 255``)."""
        c = self.config
        want_sky = "semantic" in batch and c.sky_acc_loss_mult > 0
        want_ent = c.object_acc_entropy_loss_mult > 0.0 and self.step > c.refine.stop_split_at  # background_model.stop_split_at (:386)
        fused = c.fused_loss and outputs["rgb"].is_cuda
        losses = {}

        def float_pair():  # (gt, rgb) as the torch ops consume them
            gt_img, rgb = batch["image"], outputs["rgb"]
            if gt_img.dtype == torch.uint8:
                gt_img = gt_img.float() / 255.0
            if "mask" in batch:
                gt_img, rgb = gt_img * batch["mask"], rgb * batch["mask"]
            return gt_img, rgb

        if fused:
            # one forward + one backward kernel for the three image-space terms (loss.py)
            from .loss import fused_image_losses
            l1, sky, ent = fused_image_losses(
                outputs["rgb"], batch["image"], accumulation=outputs["accumulation"] if want_sky else None,
                object_acc=outputs["object_acc"] if want_ent else None, mask=batch.get("mask"),
                sky_mask=(batch["semantic"] == 2) if want_sky else None,  # SemanticType.SKY (data/utils/data_utils.py:26-29)
                w_l1=1 - c.ssim_lambda, w_sky=c.sky_acc_loss_mult if want_sky else 0.0,
                w_entropy=c.object_acc_entropy_loss_mult if want_ent else 0.0)
            losses["Ll1"] = l1
        else:
            gt_img, rgb = float_pair()
            losses["Ll1"] = (1 - c.ssim_lambda) * torch.abs(gt_img - rgb).mean()
        if c.ssim_lambda > 0:
            gt_img, rgb = float_pair()
            simloss = 1 - ssim(gt_img.permute(2, 0, 1)[None, ...], rgb.permute(2, 0, 1)[None, ...])
            losses["simloss"] = c.ssim_lambda * simloss
        else:
            # the reference always emits the key (sgn_splatfacto.py:1085-1087); with ssim_lambda == 0 its value is 0 * simloss:
            # an exact zero that contributes no gradient, so the SSIM convolutions are not run for it
            losses["simloss"] = torch.zeros((), device=outputs["rgb"].device)
        if want_sky:
            if fused:
                losses["sky_accumulation"] = sky
            else:
                sky_mask = (batch["semantic"] == 2)
                losses["sky_accumulation"] = c.sky_acc_loss_mult * (sky_mask * outputs["accumulation"]).mean()
        if want_ent:
            if fused:
                losses["object_acc_entropy_loss"] = ent
            else:
                oa = torch.clamp(outputs["object_acc"], min=1e-5, max=1 - 1e-5)
                losses["object_acc_entropy_loss"] = c.object_acc_entropy_loss_mult * -(
                    oa * torch.log(oa) + (1.0 - oa) * torch.log(1.0 - oa)).mean()
        return losses


class _NoOptimizer:
    """refinement_after(None, ...): parameters only."""

    def relayout(self, new, changed):
        return [None] * len(new), [None] * len(new)

    def commit(self, params, changed):
        pass

    def zero_moments(self, sub: int, k: int):
        pass


class _FusedAdamAdapter(_NoOptimizer):
    """Moments live in two flat arenas: ``relayout`` allocates the arenas of the new layout, hands out views of the old
    and new ones for the sub-models that change (sgn_refine_apply writes the new ones) and block-copies the rest."""

    def __init__(self, opt):
        self.opt = opt

    def relayout(self, new, changed):
        opt = self.opt
        old_params = opt._params
        old_m, old_v, old_off = opt.rebuild(new)
        src, dst = [], []
        for i, ch in enumerate(changed):
            if ch:
                pairs = []
                for k in range(6):
                    t, o = old_params[6 * i + k], int(old_off[6 * i + k])
                    pairs.append((old_m[o:o + t.numel()].view(t.shape), old_v[o:o + t.numel()].view(t.shape)))
                src.append(pairs)
                dst.append([opt.moment_views(6 * i + k) for k in range(6)])
            else:  # same tensors, new offsets: one block copy per arena
                a, b = int(old_off[6 * i]), int(opt.offsets[6 * i])
                size = int(opt.sizes[6 * i:6 * i + 6].sum())
                opt.exp_avg[b:b + size].copy_(old_m[a:a + size])
                opt.exp_avg_sq[b:b + size].copy_(old_v[a:a + size])
                src.append(None)
                dst.append(None)
        return src, dst

    def commit(self, params, changed):
        self.opt.rebuild_pointers(params)

    def zero_moments(self, sub: int, k: int):
        m, v = self.opt.moment_views(6 * sub + k)
        m.zero_()
        v.zero_()


class _TorchAdamGroupsAdapter(_NoOptimizer):
    """The reference's form (nerfstudio ``Optimizers``): one torch.optim.Adam per group name, sub-model i's tensor at
    ``param_groups[0]["params"][i]`` (sgn_splatfacto.py:459-511 index it with ``_model_idx_in_scene_graph``)."""

    def __init__(self, groups):
        self.groups = groups
        self._new_state = {}

    def _state(self, k: int, i: int):
        opt = self.groups[PARAM_NAMES[k]]
        return opt, opt.state.get(opt.param_groups[0]["params"][i], {})

    def relayout(self, new, changed):
        src, dst = [], []
        for i, ch in enumerate(changed):
            states = [self._state(k, i)[1] for k in range(6)] if ch else []
            if not ch or not all("exp_avg" in st for st in states):  # unchanged, or never stepped: nothing to carry
                src.append(None)
                dst.append(None)
                continue
            src.append([(st["exp_avg"], st["exp_avg_sq"]) for st in states])
            dst.append([(torch.empty_like(t), torch.empty_like(t)) for t in new[i]])
            self._new_state[i] = dst[-1]
        return src, dst

    def commit(self, params, changed):
        for i, ch in enumerate(changed):
            if not ch:
                continue
            for k in range(6):
                opt, st = self._state(k, i)
                plist = opt.param_groups[0]["params"]
                opt.state.pop(plist[i], None)
                if i in self._new_state:
                    st = dict(st)
                    st["exp_avg"], st["exp_avg_sq"] = self._new_state[i][k]
                plist[i] = params[i][k]
                if st:
                    opt.state[params[i][k]] = st
        self._new_state = {}

    def zero_moments(self, sub: int, k: int):
        _, st = self._state(k, sub)
        if "exp_avg" in st:
            st["exp_avg"] = torch.zeros_like(st["exp_avg"])
            st["exp_avg_sq"] = torch.zeros_like(st["exp_avg_sq"])


def _optimizer_adapter(optimizers, num_submodels: int):
    from .optim import FusedAdam
    if optimizers is None:
        return _NoOptimizer()
    if isinstance(optimizers, FusedAdam):
        assert optimizers.num_segments == num_submodels, "FusedAdam must be built over model.optimizer_params()"
        return _FusedAdamAdapter(optimizers)
    groups = getattr(optimizers, "optimizers", optimizers)
    assert all(k in groups for k in PARAM_NAMES), f"expected Adam optimizers for the groups {PARAM_NAMES}"
    return _TorchAdamGroupsAdapter(groups)


def _gauss_window(size: int, sigma: float, device, dtype):
    x = torch.arange(size, device=device, dtype=dtype) - size // 2
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return (g / g.sum())


def ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3) as the reference configures it
    (sgn_splatfacto.py:393): separable 11-tap Gaussian, valid padding, mean over the map."""
    C = x.shape[1]
    w = _gauss_window(size, sigma, x.device, x.dtype)
    wh = w.view(1, 1, size, 1).repeat(C, 1, 1, 1)
    ww = w.view(1, 1, 1, size).repeat(C, 1, 1, 1)

    def filt(t):
        return F.conv2d(F.conv2d(t, wh, groups=C), ww, groups=C)

    K1, K2 = 0.01, 0.03
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    mu1, mu2 = filt(x), filt(y)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = filt(x * x) - mu1_sq
    s2 = filt(y * y) - mu2_sq
    s12 = filt(x * y) - mu12
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return ssim_map.flatten(2).mean(-1).mean()

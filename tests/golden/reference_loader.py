"""Imports the REFERENCE's own model source (/root/reference/street_gaussians_ns/sgn_splatfacto.py and
sgn_splatfacto_scene_graph.py) in this container so that its pure-torch methods can be executed on the CPU and their
outputs committed as golden vectors (tests/golden/make_golden_reference.py).

The reference cannot be imported as shipped: nerfstudio, gsplat, nvdiffrast, pytorch3d, kornia, pytorch_msssim and
mediapy are absent.  None of them is touched by the methods executed here -- ``refinement_after``
(with cull_gaussians / split_gaussians / dup_gaussians / dup_in_optim / remove_from_optim), ``after_train``,
``get_loss_dict`` (L1, sky accumulation, object-accumulation entropy), ``IDFT`` / ``get_fourier_features`` -- except:

  * ``gsplat._torch_impl.quat_to_rotmat`` (split_gaussians, sgn_splatfacto.py:685): restated (gsplat is not in the
    container), so the split-sample MEANS of the fixtures rest on that restatement;
  * ``pytorch_msssim.SSIM`` (get_loss_dict): stubbed to return 1, the fixtures do not hold the SSIM term.

So the missing packages are replaced by EMPTY stand-ins (names only) purely to let ``import`` succeed, and the model
objects are built without nerfstudio's constructors (``__new__`` + the attributes the methods read).  Only available
where /root/reference exists: the committed fixtures are what travels.
"""
import dataclasses
import enum
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SGN_REFERENCE_ROOT", "/root/reference")  # only mounted in the build container


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "street_gaussians_ns", "sgn_splatfacto.py"))


def _module(name: str, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behaves as a package for sub-module imports
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_module(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _unavailable(what):
    def fn(*a, **k):
        raise RuntimeError(f"{what} is not available in this container (stand-in installed by tests/golden/reference_loader.py)")
    return fn


def _quat_to_rotmat(quat):
    """gsplat 0.1.x ``_torch_impl.quat_to_rotmat`` restated: F.normalize, then the wxyz rotation-matrix formula."""
    import torch.nn.functional as F
    w, x, y, z = torch.unbind(F.normalize(quat, dim=-1), dim=-1)
    mat = torch.stack([1 - 2 * (y ** 2 + z ** 2), 2 * (x * y - w * z), 2 * (x * z + w * y),
                       2 * (x * y + w * z), 1 - 2 * (x ** 2 + z ** 2), 2 * (y * z - w * x),
                       2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x ** 2 + y ** 2)], dim=-1)
    return mat.reshape(quat.shape[:-1] + (3, 3))


class _Base(torch.nn.Module):
    """Stand-in for nerfstudio ``Model``: the one thing the executed methods use from it is ``self.device``."""

    @property
    def device(self):
        return torch.device("cpu")


@dataclasses.dataclass
class _Config:
    _target: type = dataclasses.field(default_factory=lambda: _Base)


@dataclasses.dataclass
class _CameraOptimizerConfig:
    mode: str = "off"


class _Location(enum.Enum):
    BEFORE_TRAIN_ITERATION = 1
    AFTER_TRAIN_ITERATION = 2


class _Anything:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


class Optimizers:
    """Shape of nerfstudio's ``Optimizers`` as the reference uses it: ``.optimizers[group]``."""

    def __init__(self, optimizers):
        self.optimizers = optimizers


_loaded = {}


def install_stand_ins():
    _module("gsplat")
    _module("gsplat._torch_impl", quat_to_rotmat=_quat_to_rotmat)
    _module("gsplat.project_gaussians", project_gaussians=_unavailable("gsplat.project_gaussians"))
    _module("gsplat.rasterize", rasterize_gaussians=_unavailable("gsplat.rasterize_gaussians"))
    _module("gsplat.sh", num_sh_bases=lambda d: (d + 1) ** 2, spherical_harmonics=_unavailable("gsplat.spherical_harmonics"))

    class SSIM(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, x, y):
            return torch.ones((), dtype=x.dtype)

    _module("pytorch_msssim", SSIM=SSIM)
    _module("nvdiffrast")
    _module("nvdiffrast.torch")
    _module("kornia")
    _module("mediapy")
    _module("pytorch3d")
    _module("pytorch3d.transforms", quaternion_multiply=_unavailable("pytorch3d.quaternion_multiply"))
    _module("nerfstudio")
    _module("nerfstudio.cameras")
    _module("nerfstudio.cameras.camera_optimizers", CameraOptimizer=_Anything, CameraOptimizerConfig=_CameraOptimizerConfig)
    _module("nerfstudio.cameras.cameras", Cameras=_Anything)
    _module("nerfstudio.cameras.camera_utils", quaternion_from_matrix=_unavailable("nerfstudio quaternion_from_matrix"))
    _module("nerfstudio.data")
    _module("nerfstudio.data.scene_box", OrientedBox=_Anything)
    _module("nerfstudio.engine")
    _module("nerfstudio.engine.callbacks", TrainingCallback=_Anything, TrainingCallbackAttributes=_Anything,
            TrainingCallbackLocation=_Location)
    _module("nerfstudio.engine.optimizers", Optimizers=Optimizers)
    _module("nerfstudio.models")
    _module("nerfstudio.models.base_model", Model=_Base, ModelConfig=_Config)
    _module("nerfstudio.utils")
    _module("nerfstudio.utils.colors", get_color=_unavailable("nerfstudio get_color"))
    _module("nerfstudio.utils.rich_utils", CONSOLE=types.SimpleNamespace(log=lambda *a, **k: None, print=lambda *a, **k: None))
    _module("nerfstudio.utils.colormaps")
    # the reference's own annotation / box-optimizer modules need open3d + nerfstudio internals: names only
    _module("street_gaussians_ns.data.utils.bbox_optimizers", BBoxOptimizerConfig=_Config, BBoxOptimizer=_Anything)
    _module("street_gaussians_ns.data.utils.dynamic_annotation", InterpolatedAnnotation=_Anything, Box=_Anything,
            parse_timestamp=_unavailable("parse_timestamp"))


def _load(name: str, relpath: str):
    if name in _loaded:
        return _loaded[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _loaded[name] = mod
    return mod


def load():
    """Returns (sgn_splatfacto module, sgn_splatfacto_scene_graph module) of the reference."""
    assert available(), "the reference source is not mounted here"
    pkg = _module("street_gaussians_ns")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "street_gaussians_ns")]
    for sub in ("data", "data.utils"):
        m = _module("street_gaussians_ns." + sub)
        m.__path__ = [os.path.join(REFERENCE_ROOT, "street_gaussians_ns", *sub.split("."))]
    install_stand_ins()
    _load("street_gaussians_ns.data.utils.data_utils", "street_gaussians_ns/data/utils/data_utils.py")  # the real SemanticType
    base = _load("street_gaussians_ns.sgn_splatfacto", "street_gaussians_ns/sgn_splatfacto.py")
    graph = _load("street_gaussians_ns.sgn_splatfacto_scene_graph", "street_gaussians_ns/sgn_splatfacto_scene_graph.py")
    return base, graph


def bare_model(base, params: dict, config=None, step: int = 0, num_train_data: int = 0, idx: int = 0):
    """A ``SplatfactoModel`` of the reference without nerfstudio's constructor: the attributes its training callbacks and
    ``get_loss_dict`` read (sgn_splatfacto.py:291-320 sets them in populate_modules)."""
    m = base.SplatfactoModel.__new__(base.SplatfactoModel)
    torch.nn.Module.__init__(m)
    m.config = config if config is not None else base.SplatfactoModelConfig()
    m.gauss_params = torch.nn.ParameterDict({k: torch.nn.Parameter(v.clone()) for k, v in params.items()})
    m._model_idx_in_scene_graph = idx
    m.step = step
    m.num_train_data = num_train_data
    m.xys_grad_norm = m.vis_counts = m.max_2Dsize = None
    m.refine_record_dict = {}
    m.ssim = sys.modules["pytorch_msssim"].SSIM()
    return m

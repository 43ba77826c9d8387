"""gsplat-0.1.x compatible function API backed by libsgn_raster.so (Level-1 drop-in).

The reference imports exactly five symbols from gsplat (street_gaussians_ns/sgn_splatfacto.py:11-14,
street_gaussians_ns/sgn_splatfacto_scene_graph.py:8):

    from gsplat._torch_impl import quat_to_rotmat
    from gsplat.project_gaussians import project_gaussians
    from gsplat.rasterize import rasterize_gaussians
    from gsplat.sh import num_sh_bases, spherical_harmonics

``install()`` registers this package under the name ``gsplat`` in ``sys.modules`` so the reference's
model source runs unmodified on top of the B200 kernels.
"""
import sys

from . import _torch_impl, project_gaussians as _pg, rasterize as _rs, sh as _sh
from ._torch_impl import quat_to_rotmat
from .project_gaussians import project_gaussians
from .rasterize import rasterize_gaussians
from .sh import num_sh_bases, spherical_harmonics

__version__ = "0.1.11+sgn_b200"


def install(name: str = "gsplat") -> None:
    """Make ``import gsplat`` (and its four sub-modules the reference uses) resolve to this shim."""
    me = sys.modules[__name__]
    sys.modules[name] = me
    sys.modules[name + "._torch_impl"] = _torch_impl
    sys.modules[name + ".project_gaussians"] = _pg
    sys.modules[name + ".rasterize"] = _rs
    sys.modules[name + ".sh"] = _sh

"""Runs the REFERENCE'S OWN ``SplatfactoSceneGraphModel.get_outputs`` (sgn_splatfacto_scene_graph.py:305-374 calling
sgn_splatfacto.py:793-1001) on the CPU, with the oracle's restatement of gsplat plugged into the three gsplat slots.

The reference's model code is glue around three gsplat calls (project_gaussians, spherical_harmonics,
rasterize_gaussians).  gsplat is absent from the container, but the glue itself is plain torch: here it executes
unmodified (imported by tests/golden/reference_loader.py), and each gsplat slot is served by an adapter with gsplat's
0.1.x signature built on the oracle (oracle/oracle_torch.py: project / sh_eval / blend, float32).  The result is what the
reference computes IF gsplat behaves as the oracle restates it -- which isolates the glue: scene-graph compose (Fourier
colour, object->world, concatenation order), camera -> viewmat, pre-ops, view directions and SH degree schedule, clamp,
sigmoid, the four rasterize calls (rgb+alpha, depth, objects-only, background-only), the post-ops, the early-outs and
the side-effect attributes.  tests/test_reference_glue.py compares it with the oracle's OWN restatement of that glue
(oracle_torch.render + oracle_c.post_ops) on the same scene: any difference is a glue bug in the oracle (and therefore
in what the CUDA path is held to).

Restated stand-ins used by the executed reference code (both tiny, both documented library functions):
``nerfstudio.cameras.camera_utils.quaternion_from_matrix`` and ``pytorch3d.transforms.quaternion_multiply``
(raw Hamilton product, then standardised to a non-negative real part).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import reference_loader as rl  # noqa: E402
from oracle import oracle_torch as ot  # noqa: E402

PARAMS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")


# ---------------------------------------------------------------------------------------------------------------
# gsplat 0.1.x slots served by the oracle
# ---------------------------------------------------------------------------------------------------------------
class _View:
    """What oracle_torch.project / blend read from a camera, built from gsplat's call arguments."""

    def __init__(self, viewmat, fx, fy, cx, cy, H, W):
        self._vm = None if viewmat is None else viewmat.detach().cpu().numpy().astype(np.float32)
        self.fx, self.fy, self.cx, self.cy, self.height, self.width = fx, fy, cx, cy, int(H), int(W)

    def viewmat(self):
        return self._vm

    def fov_limits(self):
        tan_x = np.float32(0.5 * self.width / self.fx)
        tan_y = np.float32(0.5 * self.height / self.fy)
        return float(np.float32(1.3) * tan_x), float(np.float32(1.3) * tan_y)


def project_gaussians(means3d, scales, glob_scale, quats, viewmat, fx, fy, cx, cy, img_height, img_width, block_width,
                      clip_thresh=0.01):
    assert glob_scale == 1
    cat = dict(means=means3d, quats=quats, scales=scales)
    pr = ot.project(cat, _View(viewmat, fx, fy, cx, cy, img_height, img_width), block_width, clip_thresh, scales_are_linear=True)
    n = means3d.shape[0]
    return pr["xys"], pr["depths"], pr["radii"], pr["conics"], torch.zeros(n), pr["num_tiles_hit"], torch.zeros(n, 6)


def spherical_harmonics(degrees_to_use, viewdirs, coeffs):
    return ot.sh_eval(degrees_to_use, viewdirs, coeffs)


def bin_and_sort(xys, depths, radii, H, W, bw):
    """gsplat map_gaussian_to_intersects + sort + get_tile_bin_edges: 64-bit keys
    (tile << 32 | float bits of depth), entries of a Gaussian emitted over its tile AABB."""
    xy = xys.detach().cpu().numpy().astype(np.float32)
    d = depths.detach().cpu().numpy().astype(np.float32)
    r = radii.detach().cpu().numpy()
    tiles_x, tiles_y = (W + bw - 1) // bw, (H + bw - 1) // bw
    keys, ids = [], []
    for g in np.nonzero(r > 0)[0]:
        tc = xy[g] / np.float32(bw)
        tr = np.float32(r[g]) / np.float32(bw)
        x0 = min(max(0, int(np.trunc(tc[0] - tr))), tiles_x)
        x1 = min(max(0, int(np.trunc(tc[0] + tr + 1))), tiles_x)
        y0 = min(max(0, int(np.trunc(tc[1] - tr))), tiles_y)
        y1 = min(max(0, int(np.trunc(tc[1] + tr + 1))), tiles_y)
        bits = int(d[g:g + 1].view(np.int32)[0])
        for ty in range(y0, y1):
            for tx in range(x0, x1):
                keys.append(((ty * tiles_x + tx) << 32) | (bits & 0xFFFFFFFF))
                ids.append(g)
    keys = np.asarray(keys, np.int64)
    ids = np.asarray(ids, np.int64)
    order = np.argsort(keys, kind="stable")
    keys, ids = keys[order], ids[order]
    bins = np.zeros((tiles_x * tiles_y, 2), np.int64)
    tile = keys >> 32
    for t in range(tiles_x * tiles_y):
        sel = np.nonzero(tile == t)[0]
        if len(sel):
            bins[t] = (sel[0], sel[-1] + 1)
    return ids, bins


def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, block_width,
                        background=None, return_alpha=False):
    H, W = int(img_height), int(img_width)
    ids, bins = bin_and_sort(xys, depths, radii, H, W, block_width)
    img, alpha = ot.blend(_View(None, 1.0, 1.0, 0.0, 0.0, H, W), ids, bins, xys, conics, colors, opacity[:, 0], block_width,
                          alpha_clamp=0.999)
    if background is not None:
        img = img + (1.0 - alpha)[..., None] * background
    return (img, alpha) if return_alpha else img


def quaternion_multiply(a, b):
    """pytorch3d.transforms.quaternion_multiply: Hamilton product (real part first), standardised to real >= 0."""
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    q = torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)
    return torch.where(q[..., 0:1] < 0, -q, q)


def quaternion_from_matrix(matrix):
    from street_gaussians_ns_b200.scene import quaternion_from_matrix as qfm  # restatement of nerfstudio's (Hartley eigen form)
    return qfm(np.asarray(matrix))


# ---------------------------------------------------------------------------------------------------------------
# a reference scene-graph model for a synthetic Frame, without nerfstudio's constructors
# ---------------------------------------------------------------------------------------------------------------
class _Annotations:
    """The three things get_outputs reads from ``InterpolatedAnnotation``: boxes at a timestamp, annotated frame names,
    every track's frame list."""

    def __init__(self, annos, frames):
        self.annos, self.all_names, self.objects_frames = annos, set(), frames

    def __getitem__(self, timestamp):
        return self.annos


def build_reference_model(frame, training: bool, step: int = 30000, sky=None, num_frames: int = 85):
    """``frame``: street_gaussians_ns_b200.scene.Frame (background first, then actors with rot / center / name).
    Returns (model, camera) of the reference's classes."""
    base, graph = rl.load()
    base.project_gaussians, base.spherical_harmonics, base.rasterize_gaussians = project_gaussians, spherical_harmonics, rasterize_gaussians
    graph.spherical_harmonics = spherical_harmonics
    graph.quaternion_multiply, graph.quaternion_from_matrix = quaternion_multiply, quaternion_from_matrix
    graph.parse_timestamp = lambda t: str(t)
    cls = graph.SplatfactoSceneGraphModel
    m = cls.__new__(cls)
    torch.nn.Module.__init__(m)
    m.config = graph.SplatfactoSceneGraphModelConfig()
    m.config.use_sky_sphere = sky is not None
    m.gauss_params = torch.nn.ParameterDict()
    for k in PARAMS:                      # what populate_modules leaves behind (scene graph :53-57): plain attributes
        setattr(m, k, None)
    m._xys = m._radii = m._depths = m._conics = m._num_tiles_hit = m._last_size = None
    m.all_models = torch.nn.ModuleDict()
    annos = []
    for i, seg in enumerate(frame.segments):
        params = {k: getattr(seg.params, k).detach().clone() for k in PARAMS}
        sub_cfg = base.SplatfactoModelConfig(use_sky_sphere=False, fourier_features_dim=int(params["features_dc"].shape[1]))
        sub = rl.bare_model(base, params, sub_cfg, step=step, idx=i)
        sub.xys = sub.depths = sub.radii = sub.conics = sub.num_tiles_hit = sub.last_size = None
        name = "background" if i == 0 else seg.name
        m.all_models[name] = sub
        if i > 0:
            track = name[len("object_"):]
            annos.append(types.SimpleNamespace(trackId=track, frame=int(frame.camera.time), center=np.asarray(seg.center, np.float64),
                                               rot=np.asarray(seg.rot, np.float64)))
    # the reference sizes features_dc by config.fourier_features_dim (sgn_splatfacto.py:272-289): keep config and data consistent
    m.config.fourier_features_dim = max([int(s.params.features_dc.shape[1]) for s in frame.segments[1:]] or [1])
    m.object_annos = _Annotations(annos, {a.trackId: list(range(num_frames)) for a in annos})
    m.bbox_optimizer = types.SimpleNamespace(apply_to_bbox=lambda anno: None)
    m.visible_model_names = list(m.all_models.keys())
    m.back_color = torch.zeros(3)
    m.crop_box = None
    m.step = step
    m.ssim = sys.modules["pytorch_msssim"].SSIM()
    if sky is not None:
        m.env_map = lambda camera, training: sky
    m.train(training)
    c = frame.camera
    cam = sys.modules["nerfstudio.cameras.cameras"].Cameras()
    cam.shape = (1,)
    cam.camera_to_worlds = torch.from_numpy(np.asarray(c.c2w, np.float32))[None]
    cam.fx, cam.fy = torch.tensor([[np.float32(c.fx)]]), torch.tensor([[np.float32(c.fy)]])
    cam.cx, cam.cy = torch.tensor([[np.float32(c.cx)]]), torch.tensor([[np.float32(c.cy)]])
    cam.width, cam.height = torch.tensor([[c.width]]), torch.tensor([[c.height]])
    cam.times = torch.tensor([[float(c.time)]], dtype=torch.float64)
    cam.rescale_output_resolution = lambda s: None
    return m, cam

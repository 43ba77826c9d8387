"""ctypes front-end of the C oracle (oracle/sgn_oracle.c).  TEST INFRASTRUCTURE -- never imported
by the product package.  Mirrors what ``SplatfactoSceneGraphModel.get_outputs`` computes
(street_gaussians_ns/sgn_splatfacto_scene_graph.py:305-374) on CPU tensors, including the extra
objects-only / background-only accumulation renders and the reference's post-ops
(street_gaussians_ns/sgn_splatfacto.py:968-996).

PARITY UNPINNED (see sgn_oracle.c header): no reference golden vectors exist for this path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsgn_oracle.so")
MAX_F = 8


class OracleSegment(C.Structure):
    _fields_ = [
        ("row0", C.c_int32), ("count", C.c_int32), ("F", C.c_int32), ("cls", C.c_int32),
        ("has_pose", C.c_int32), ("pad0", C.c_int32),
        ("R", C.c_float * 9), ("t", C.c_float * 3), ("q", C.c_float * 4), ("idft", C.c_float * MAX_F),
        ("means", C.c_void_p), ("scales", C.c_void_p), ("quats", C.c_void_p),
        ("features_dc", C.c_void_p), ("features_rest", C.c_void_p), ("opacities", C.c_void_p),
    ]


class OracleCamera(C.Structure):
    _fields_ = [
        ("viewmat", C.c_float * 12),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("width", C.c_int32), ("height", C.c_int32),
        ("cam_pos", C.c_float * 3),
        ("limx", C.c_float), ("limy", C.c_float),
        ("clip_thresh", C.c_float),
        ("block_width", C.c_int32),
        ("sh_degree", C.c_int32), ("sh_degree_to_use", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "sgn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libsgn_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.sgn_oracle_bin_sort.restype = C.c_int64
        _lib.sgn_expf_spec.restype = C.c_float
        _lib.sgn_expf_spec.argtypes = [C.c_float]
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def expf_spec(x: np.ndarray) -> np.ndarray:
    L = lib()
    return np.array([L.sgn_expf_spec(float(v)) for v in np.asarray(x, dtype=np.float32).reshape(-1)],
                    dtype=np.float32).reshape(np.shape(x))


def _camera_struct(cam, sh_degree: int, sh_degree_to_use: int, block_width: int, clip: float) -> OracleCamera:
    oc = OracleCamera()
    vm = cam.viewmat().reshape(-1)
    for i in range(12):
        oc.viewmat[i] = float(vm[i])
    oc.fx, oc.fy, oc.cx, oc.cy = cam.fx, cam.fy, cam.cx, cam.cy
    oc.width, oc.height = cam.width, cam.height
    cp = cam.cam_pos()
    for i in range(3):
        oc.cam_pos[i] = float(cp[i])
    oc.limx, oc.limy = cam.fov_limits()
    oc.clip_thresh = clip
    oc.block_width = block_width
    oc.sh_degree = sh_degree
    oc.sh_degree_to_use = sh_degree_to_use
    return oc


@dataclass
class OracleForward:
    """Everything the forward produced (numpy, dense over the concatenated row space)."""

    N: int
    M: int
    xys: np.ndarray
    depths: np.ndarray
    radii: np.ndarray
    conics: np.ndarray
    num_tiles_hit: np.ndarray
    tile_bbox: np.ndarray
    rgbs: np.ndarray
    rgb_pre: np.ndarray
    opac: np.ndarray
    cls: np.ndarray
    sorted_ids: np.ndarray
    tile_bins: np.ndarray
    # raw blend outputs
    img: np.ndarray  # [H,W,4] rgb + depth channel, premultiplied, background 0
    final_T: np.ndarray
    final_idx: np.ndarray
    fragile: np.ndarray
    obj_T: Optional[np.ndarray] = None
    obj_idx: Optional[np.ndarray] = None
    bg_T: Optional[np.ndarray] = None
    bg_idx: Optional[np.ndarray] = None
    fragile_obj: Optional[np.ndarray] = None
    fragile_bg: Optional[np.ndarray] = None
    keep: list = field(default_factory=list)  # keeps segment arrays alive


class Oracle:
    """Stateless helper bound to (frame, render settings)."""

    def __init__(self, frame, sh_degree: int = 3, sh_degree_to_use: Optional[int] = None, block_width: int = 16,
                 clip_thresh: float = 0.01, alpha_clamp_fwd: float = 0.999, alpha_clamp_bwd: float = 0.99,
                 margin: float = 1e-4):
        self.frame = frame
        self.sh_degree = sh_degree
        self.sh_degree_to_use = sh_degree if sh_degree_to_use is None else sh_degree_to_use
        self.bw = block_width
        self.clip = clip_thresh
        self.clamp_fwd = alpha_clamp_fwd
        self.clamp_bwd = alpha_clamp_bwd
        self.margin = margin
        self.L = lib()
        self.cam = _camera_struct(frame.camera, sh_degree, self.sh_degree_to_use, block_width, clip_thresh)
        self._keep = []
        self.segs = (OracleSegment * len(frame.segments))()
        row = 0
        for i, s in enumerate(frame.segments):
            p = s.params
            p.validate(sh_degree)
            arrs = [t.detach().cpu().contiguous().numpy() for t in p.tensors()]
            self._keep.append(arrs)
            sg = self.segs[i]
            sg.row0, sg.count, sg.F, sg.cls, sg.has_pose = row, p.num_points, p.fourier_dim, s.cls, int(s.has_pose)
            R, t, q = s.pose_f32()
            for k in range(9):
                sg.R[k] = float(R[k])
            for k in range(3):
                sg.t[k] = float(t[k])
            for k in range(4):
                sg.q[k] = float(q[k])
            idft = s.idft_f32()
            for k in range(MAX_F):
                sg.idft[k] = float(idft[k])
            (sg.means, sg.scales, sg.quats, sg.features_dc, sg.features_rest, sg.opacities) = [
                a.ctypes.data for a in arrs
            ]
            row += p.num_points
        self.N = row
        self.cls = np.concatenate(
            [np.full(s.params.num_points, s.cls, dtype=np.int32) for s in frame.segments]
        ) if frame.segments else np.zeros(0, np.int32)

    # ---------------------------------------------------------------- forward
    def project(self):
        N = self.N
        out = dict(
            xys=np.zeros((N, 2), np.float32), depths=np.zeros(N, np.float32), radii=np.zeros(N, np.int32),
            conics=np.zeros((N, 3), np.float32), num_tiles_hit=np.zeros(N, np.int32),
            rgbs=np.zeros((N, 3), np.float32), rgb_pre=np.zeros((N, 3), np.float32),
            opac=np.zeros(N, np.float32), tile_bbox=np.zeros((N, 4), np.int32),
        )
        rc = self.L.sgn_oracle_project(
            self.segs, len(self.frame.segments), C.byref(self.cam), _p(out["xys"]), _p(out["depths"]),
            _p(out["radii"]), _p(out["conics"]), _p(out["num_tiles_hit"]), _p(out["rgbs"]),
            _p(out["rgb_pre"]), _p(out["opac"]), _p(out["tile_bbox"]))
        assert rc == 0
        return out

    def bin_sort(self, pr):
        cam = self.frame.camera
        tiles = ((cam.width + self.bw - 1) // self.bw) * ((cam.height + self.bw - 1) // self.bw)
        M = int(pr["num_tiles_hit"].astype(np.int64).sum())
        sorted_ids = np.zeros(max(M, 1), np.int32)
        tile_bins = np.zeros((tiles, 2), np.int32)
        got = self.L.sgn_oracle_bin_sort(
            C.c_int(self.N), _p(pr["xys"]), _p(pr["depths"]), _p(pr["radii"]), _p(pr["num_tiles_hit"]),
            C.c_int(cam.width), C.c_int(cam.height), C.c_int(self.bw), _p(sorted_ids), _p(tile_bins))
        assert got == M, f"bin_sort returned {got}, expected {M}"
        return M, sorted_ids[:M] if M else sorted_ids[:0], tile_bins

    def entry_any_valid(self, fw) -> np.ndarray:
        """uint8[M]: can any pixel centre of the entry's tile accept it (alpha >= 1/255)?"""
        cam = self.frame.camera
        out = np.zeros(max(fw.M, 1), np.uint8)
        ids = fw.sorted_ids if fw.sorted_ids.size else np.zeros(1, np.int32)
        rc = self.L.sgn_oracle_entry_any_valid(
            C.c_int(cam.width), C.c_int(cam.height), C.c_int(self.bw), _p(ids), _p(fw.tile_bins), _p(fw.xys),
            _p(fw.conics), _p(fw.opac), _p(out))
        assert rc == 0
        return out[: fw.M]

    def blend(self, pr, sorted_ids, tile_bins, colors: np.ndarray, cls_filter: int = -1,
              background: Optional[np.ndarray] = None):
        cam = self.frame.camera
        H, W = cam.height, cam.width
        Cc = colors.shape[1] if colors is not None and colors.size else 0
        colors = np.ascontiguousarray(colors, np.float32) if Cc else np.zeros((self.N, 0), np.float32)
        bgc = np.zeros(max(Cc, 1), np.float32) if background is None else np.ascontiguousarray(background, np.float32)
        img = np.zeros((H, W, Cc), np.float32)
        fT = np.zeros((H, W), np.float32)
        fi = np.zeros((H, W), np.int32)
        frag = np.zeros((H, W), np.uint8)
        ids = sorted_ids if sorted_ids.size else np.zeros(1, np.int32)
        rc = self.L.sgn_oracle_blend_fwd(
            C.c_int(W), C.c_int(H), C.c_int(self.bw), C.c_int(Cc), _p(ids), _p(tile_bins), _p(pr["xys"]),
            _p(pr["conics"]), _p(colors if Cc else np.zeros(1, np.float32)), _p(pr["opac"]), _p(bgc),
            C.c_float(self.clamp_fwd), _p(self.cls), C.c_int(cls_filter), _p(img if Cc else np.zeros(1, np.float32)),
            _p(fT), _p(fi), _p(frag), C.c_float(self.margin))
        assert rc == 0
        return img, fT, fi, frag

    def forward(self, class_renders: bool = True) -> OracleForward:
        pr = self.project()
        M, sorted_ids, tile_bins = self.bin_sort(pr)
        colors4 = np.concatenate([pr["rgbs"], pr["depths"][:, None]], axis=1)
        img, fT, fi, frag = self.blend(pr, sorted_ids, tile_bins, colors4)
        fw = OracleForward(
            N=self.N, M=M, xys=pr["xys"], depths=pr["depths"], radii=pr["radii"], conics=pr["conics"],
            num_tiles_hit=pr["num_tiles_hit"], tile_bbox=pr["tile_bbox"], rgbs=pr["rgbs"], rgb_pre=pr["rgb_pre"],
            opac=pr["opac"], cls=self.cls, sorted_ids=sorted_ids, tile_bins=tile_bins, img=img, final_T=fT,
            final_idx=fi, fragile=frag)
        if class_renders:
            _, fw.obj_T, fw.obj_idx, fw.fragile_obj = self.blend(pr, sorted_ids, tile_bins, None, cls_filter=1)
            _, fw.bg_T, fw.bg_idx, fw.fragile_bg = self.blend(pr, sorted_ids, tile_bins, None, cls_filter=0)
        return fw

    # ---------------------------------------------------------------- backward
    def blend_bwd(self, fw: OracleForward, colors: np.ndarray, fT, fi, v_img, v_alpha, cls_filter=-1, acc=None):
        cam = self.frame.camera
        H, W = cam.height, cam.width
        Cc = colors.shape[1] if colors is not None and colors.size else 0
        N = self.N
        if acc is None:
            acc = dict(v_xy=np.zeros((N, 2), np.float64), v_conic=np.zeros((N, 3), np.float64),
                       v_colors=np.zeros((N, 4), np.float64), v_opac=np.zeros(N, np.float64))
        vc = np.zeros((N, max(Cc, 1)), np.float64)
        colors = np.ascontiguousarray(colors, np.float32) if Cc else np.zeros(1, np.float32)
        bgc = np.zeros(max(Cc, 1), np.float32)
        v_img = np.ascontiguousarray(v_img, np.float32) if Cc else np.zeros(1, np.float32)
        v_alpha = np.ascontiguousarray(v_alpha, np.float32)
        ids = fw.sorted_ids if fw.sorted_ids.size else np.zeros(1, np.int32)
        rc = self.L.sgn_oracle_blend_bwd(
            C.c_int(W), C.c_int(H), C.c_int(self.bw), C.c_int(Cc), _p(ids), _p(fw.tile_bins), _p(fw.xys),
            _p(fw.conics), _p(colors), _p(fw.opac), _p(bgc), C.c_float(self.clamp_bwd), _p(self.cls),
            C.c_int(cls_filter), _p(np.ascontiguousarray(fT)), _p(np.ascontiguousarray(fi)), _p(v_img),
            _p(v_alpha), _p(acc["v_xy"]), _p(acc["v_conic"]), _p(vc), _p(acc["v_opac"]))
        assert rc == 0
        if Cc:
            acc["v_colors"][:, :Cc] += vc[:, :Cc]
        return acc

    def backward(self, fw: OracleForward, v_img4: np.ndarray, v_alpha: np.ndarray,
                 v_obj_alpha: Optional[np.ndarray] = None, v_bg_alpha: Optional[np.ndarray] = None):
        """Cotangents w.r.t. the RAW blend outputs (img[H,W,4] = rgb+depth channel, alpha = 1-final_T,
        object/background accumulation) -> per-segment parameter gradients + per-Gaussian raster grads."""
        colors4 = np.concatenate([fw.rgbs, fw.depths[:, None]], axis=1)
        acc = self.blend_bwd(fw, colors4, fw.final_T, fw.final_idx, v_img4, v_alpha)
        if v_obj_alpha is not None:
            self.blend_bwd(fw, None, fw.obj_T, fw.obj_idx, None, v_obj_alpha, cls_filter=1, acc=acc)
        if v_bg_alpha is not None:
            self.blend_bwd(fw, None, fw.bg_T, fw.bg_idx, None, v_bg_alpha, cls_filter=0, acc=acc)
        v_xy = acc["v_xy"].astype(np.float32)
        v_conic = acc["v_conic"].astype(np.float32)
        v_rgb = np.ascontiguousarray(acc["v_colors"][:, :3].astype(np.float32))
        v_depth = np.ascontiguousarray(acc["v_colors"][:, 3].astype(np.float32))
        v_opac = acc["v_opac"].astype(np.float32)
        grads = self.project_bwd(fw, v_xy, v_depth, v_conic, v_rgb, v_opac)
        raster = dict(v_xy=v_xy, v_conic=v_conic, v_rgb=v_rgb, v_depth=v_depth, v_opac=v_opac)
        return grads, raster

    def project_bwd(self, fw: OracleForward, v_xy, v_depth, v_conic, v_rgb, v_opac):
        nseg = len(self.frame.segments)
        K = (self.sh_degree + 1) ** 2
        outs: List[Dict[str, np.ndarray]] = []
        ptrs = {k: (C.c_void_p * nseg)() for k in ("means", "scales", "quats", "dc", "rest", "opac")}
        for i, s in enumerate(self.frame.segments):
            n, F = s.params.num_points, s.params.fourier_dim
            g = dict(means=np.zeros((n, 3), np.float32), scales=np.zeros((n, 3), np.float32),
                     quats=np.zeros((n, 4), np.float32), dc=np.zeros((n, F, 3), np.float32),
                     rest=np.zeros((n, K - 1, 3), np.float32), opac=np.zeros((n, 1), np.float32))
            outs.append(g)
            for k in ptrs:
                ptrs[k][i] = g[k].ctypes.data
        rc = self.L.sgn_oracle_project_bwd(
            self.segs, C.c_int(nseg), C.byref(self.cam), _p(fw.radii), _p(fw.rgb_pre),
            _p(np.ascontiguousarray(v_xy, np.float32)), _p(np.ascontiguousarray(v_depth, np.float32)),
            _p(np.ascontiguousarray(v_conic, np.float32)), _p(np.ascontiguousarray(v_rgb, np.float32)),
            _p(np.ascontiguousarray(v_opac, np.float32)),
            ptrs["means"], ptrs["scales"], ptrs["quats"], ptrs["dc"], ptrs["rest"], ptrs["opac"])
        assert rc == 0
        return [
            dict(means=g["means"], scales=g["scales"], quats=g["quats"], features_dc=g["dc"],
                 features_rest=g["rest"], opacities=g["opac"]) for g in outs
        ]


# ---------------------------------------------------------------------------------------------
# reference post-ops (street_gaussians_ns/sgn_splatfacto.py:968-996; scene graph :364-366)
# ---------------------------------------------------------------------------------------------
def post_ops(img4: torch.Tensor, alpha: torch.Tensor, sky: Optional[torch.Tensor], training: bool):
    """rgb = clamp(rgb, max=1); rgb = rgb*alpha + sky*(1-alpha) [premultiplied rgb multiplied by
    alpha AGAIN, Appendix B.1]; eval: clamp(0,1); depth = where(alpha>1e-3, d/alpha, 10)."""
    a = alpha[..., None]
    rgb = torch.clamp(img4[..., :3], max=1.0)
    if sky is not None:
        rgb = rgb * a + sky * (1 - a)
    if not training:
        rgb = rgb.clamp(0.0, 1.0)
    # same values as torch.where(alpha > 1e-3, depth_im / alpha, 10); the guarded denominator only keeps
    # autograd from producing 0/0 in the unselected branch
    safe = torch.where(a > 1e-3, a, torch.ones_like(a))
    depth = torch.where(a > 1e-3, img4[..., 3:4] / safe, torch.full_like(a, 10.0))
    return rgb, a, depth


def num_threads() -> int:
    return int(lib().sgn_oracle_num_threads())

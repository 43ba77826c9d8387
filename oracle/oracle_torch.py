"""Pure-PyTorch restatement of the hot path with AUTOGRAD backward (float32 or float64).

TEST INFRASTRUCTURE -- never imported by the product.  Independent of oracle/sgn_oracle.c: it is
written from the same sources (street_gaussians_ns/sgn_splatfacto.py:822-873,916-1001,
street_gaussians_ns/sgn_splatfacto_scene_graph.py:239-247,404-433) but
vectorised, and its gradients come from autograd instead of hand-derived VJPs.  It pins the C
oracle (and through it the CUDA path): forward values, the backward formulas (in the
consistent-clamp mode) and finite differences in float64.

PARITY UNPINNED: gsplat 0.1.x is absent and the reference has no golden vectors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def expf_spec(x: torch.Tensor) -> torch.Tensor:
    """The fixed-sequence exp of the exact section (see oracle/sgn_oracle.c: sgn_expf_spec)."""
    one = lambda v: torch.tensor(v, dtype=x.dtype)
    x = x.clamp(-80.0, 80.0)
    n = torch.round(x * one(1.44269504))
    r = x - n * one(0.693145752)
    r = r - n * one(1.42860677e-6)
    p = one(1.98412698e-4)
    for c in (1.38888889e-3, 8.33333333e-3, 4.16666667e-2, 1.66666667e-1, 0.5, 1.0, 1.0):
        p = p * r + one(c)
    # torch.ldexp's autograd mishandles negative integer exponents: build the exact power of two
    # without grad and multiply
    with torch.no_grad():
        scale = torch.ldexp(torch.ones_like(p), n.to(torch.int32))
    return p * scale


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.quaternion_raw_multiply (Hamilton, real first)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack([ow, ox, oy, oz], -1)


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = q.unbind(-1)
    return torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        -1,
    ).reshape(*q.shape[:-1], 3, 3)


def sh_eval(deg: int, dirs: torch.Tensor, coeffs: torch.Tensor) -> torch.Tensor:
    """gsplat spherical_harmonics (Appendix A.7).  dirs[N,3] unit, coeffs[N,K,3] -> [N,3]."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    out = SH_C0 * coeffs[:, 0]
    if deg < 1:
        return out
    out = out - SH_C1 * y * coeffs[:, 1] + SH_C1 * z * coeffs[:, 2] - SH_C1 * x * coeffs[:, 3]
    if deg < 2:
        return out
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    out = (out + SH_C2[0] * xy * coeffs[:, 4] + SH_C2[1] * yz * coeffs[:, 5]
           + SH_C2[2] * (2 * zz - xx - yy) * coeffs[:, 6] + SH_C2[3] * xz * coeffs[:, 7]
           + SH_C2[4] * (xx - yy) * coeffs[:, 8])
    if deg < 3:
        return out
    out = (out + SH_C3[0] * y * (3 * xx - yy) * coeffs[:, 9] + SH_C3[1] * xy * z * coeffs[:, 10]
           + SH_C3[2] * y * (4 * zz - xx - yy) * coeffs[:, 11]
           + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * coeffs[:, 12]
           + SH_C3[4] * x * (4 * zz - xx - yy) * coeffs[:, 13] + SH_C3[5] * z * (xx - yy) * coeffs[:, 14]
           + SH_C3[6] * x * (xx - 3 * yy) * coeffs[:, 15])
    return out


class _ClampMaxST(torch.autograd.Function):
    """min(x, c) whose backward passes the gradient straight through (gsplat's rasterize backward
    does not zero v_sigma when the alpha clamp is active, Appendix A.6)."""

    @staticmethod
    def forward(ctx, x, c):
        return torch.clamp(x, max=c)

    @staticmethod
    def backward(ctx, g):
        return g, None


def compose(frame, dtype=torch.float64, requires_grad: bool = True):
    """Leaf copies of every segment's parameters in ``dtype`` + the composed world-space tensors
    (scene graph :332-360).  Returns (leaves, dict of concatenated tensors)."""
    leaves: List[Dict[str, torch.Tensor]] = []
    means, quats, dcs, rests, scales, opacs, cls = [], [], [], [], [], [], []
    for s in frame.segments:
        p = s.params
        lf = {k: getattr(p, k).detach().cpu().to(dtype).clone().requires_grad_(requires_grad)
              for k in ("means", "scales", "quats", "features_dc", "features_rest", "opacities")}
        leaves.append(lf)
        idft = torch.from_numpy(s.idft_f32()[: p.fourier_dim].copy()).to(dtype)
        fdc = torch.sum(lf["features_dc"] * idft[None, :, None], dim=1, keepdim=True)
        if s.has_pose:
            R, t, q = s.pose_f32()
            R = torch.from_numpy(R.reshape(3, 3).copy()).to(dtype)
            t = torch.from_numpy(t.copy()).to(dtype)
            q = torch.from_numpy(q.copy()).to(dtype)
            m = lf["means"]
            mw = torch.stack(
                [R[0, 0] * m[:, 0] + R[0, 1] * m[:, 1] + R[0, 2] * m[:, 2] + t[0],
                 R[1, 0] * m[:, 0] + R[1, 1] * m[:, 1] + R[1, 2] * m[:, 2] + t[1],
                 R[2, 0] * m[:, 0] + R[2, 1] * m[:, 1] + R[2, 2] * m[:, 2] + t[2]], -1)
            qw = quat_mul(q[None, :].expand_as(lf["quats"]), lf["quats"])
        else:
            mw, qw = lf["means"], lf["quats"]
        means.append(mw); quats.append(qw); dcs.append(fdc); rests.append(lf["features_rest"])
        scales.append(lf["scales"]); opacs.append(lf["opacities"])
        cls.append(torch.full((p.num_points,), s.cls, dtype=torch.int32))
    cat = dict(means=torch.cat(means), quats=torch.cat(quats), features_dc=torch.cat(dcs),
               features_rest=torch.cat(rests), scales=torch.cat(scales), opacities=torch.cat(opacs),
               cls=torch.cat(cls))
    return leaves, cat


def project(cat, camera, block_width=16, clip_thresh=0.01, use_spec_exp=True, scales_are_linear=False):
    """gsplat project_gaussians on the composed tensors (Appendix A.1-A.4).  ``scales_are_linear``: cat["scales"] already
    went through exp, as in the reference's call (sgn_splatfacto.py:857-862) -- used when this function serves the
    gsplat slot under the reference's own glue (tests/golden/reference_glue.py)."""
    dtype = cat["means"].dtype
    W = torch.from_numpy(camera.viewmat().copy()).to(dtype)  # [3,4]
    fx, fy, cx, cy = camera.fx, camera.fy, camera.cx, camera.cy
    limx, limy = camera.fov_limits()
    m = cat["means"]
    pv = torch.stack(
        [W[r, 0] * m[:, 0] + W[r, 1] * m[:, 1] + W[r, 2] * m[:, 2] + W[r, 3] for r in range(3)], -1)
    z = pv[:, 2]
    not_clipped = z > clip_thresh
    q = cat["quats"]
    qn = q / q.norm(dim=-1, keepdim=True)
    if scales_are_linear:
        s = cat["scales"]
    else:
        s = expf_spec(cat["scales"]) if use_spec_exp else torch.exp(cat["scales"])
    Rg = quat_to_rotmat(qn)
    M = Rg * s[:, None, :]
    S = M @ M.transpose(1, 2)
    zs = torch.where(not_clipped, z, torch.ones_like(z))
    rz = 1.0 / zs
    rz2 = rz * rz
    ux = torch.clamp(pv[:, 0] / zs, -limx, limx)
    uy = torch.clamp(pv[:, 1] / zs, -limy, limy)
    tx, ty = zs * ux, zs * uy
    zero = torch.zeros_like(z)
    J = torch.stack([fx * rz, zero, -fx * tx * rz2, zero, fy * rz, -fy * ty * rz2], -1).reshape(-1, 2, 3)
    T = J @ W[:, :3]
    cov = T @ S @ T.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    ok = not_clipped & (det != 0)
    dets = torch.where(ok, det, torch.ones_like(det))
    conics = torch.stack([c / dets, -b / dets, a / dets], -1)
    bm = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp(bm * bm - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(bm + disc, bm - disc)))
    rw = 1.0 / (zs + 1e-6)
    xys = torch.stack([pv[:, 0] * rw * fx + cx, pv[:, 1] * rw * fy + cy], -1)
    tiles_x = (camera.width + block_width - 1) // block_width
    tiles_y = (camera.height + block_width - 1) // block_width
    tc = xys.detach() / block_width
    tr = (radius.detach() / block_width)[:, None]
    tmin = torch.trunc(tc - tr).clamp(min=0)
    tmax = torch.trunc(tc + tr + 1).clamp(min=0)
    tmin = torch.minimum(tmin, torch.tensor([tiles_x, tiles_y], dtype=dtype))
    tmax = torch.minimum(tmax, torch.tensor([tiles_x, tiles_y], dtype=dtype))
    area = ((tmax[:, 0] - tmin[:, 0]) * (tmax[:, 1] - tmin[:, 1])).to(torch.int64)
    vis = ok & (area > 0)
    radii = torch.where(vis, radius.detach(), torch.zeros_like(radius)).to(torch.int32)
    return dict(xys=torch.where(vis[:, None], xys, torch.zeros_like(xys)),
                depths=torch.where(vis, z, torch.zeros_like(z)),
                radii=radii, conics=torch.where(ok[:, None], conics, torch.zeros_like(conics)),
                num_tiles_hit=torch.where(vis, area, torch.zeros_like(area)).to(torch.int32), visible=vis)


def colours(cat, camera, sh_degree: int, sh_degree_to_use: int):
    """SH colour + clamp and sigmoid opacity (street_gaussians_ns/sgn_splatfacto.py:933-949)."""
    dtype = cat["means"].dtype
    colors = torch.cat((cat["features_dc"], cat["features_rest"]), dim=1)
    if sh_degree > 0:
        cam_pos = torch.from_numpy(camera.cam_pos().copy()).to(dtype)
        viewdirs = cat["means"].detach() - cam_pos
        viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
        rgbs = torch.clamp(sh_eval(sh_degree_to_use, viewdirs, colors) + 0.5, min=0.0)
    else:
        rgbs = torch.sigmoid(colors[:, 0, :])
    return rgbs, torch.sigmoid(cat["opacities"])[:, 0]


def blend(camera, sorted_ids: np.ndarray, tile_bins: np.ndarray, xys, conics, colors, opac,
          block_width: int = 16, alpha_clamp: float = 0.999, cls: Optional[torch.Tensor] = None,
          cls_filter: int = -1):
    """Differentiable front-to-back compositing over given per-tile sorted lists (Appendix A.6).
    Returns (img[H,W,C], alpha[H,W]).  Skip / early-termination masks are recomputed here."""
    H, W = camera.height, camera.width
    dtype = xys.dtype
    C = colors.shape[1]
    img = torch.zeros(H, W, C, dtype=dtype)
    alpha_out = torch.zeros(H, W, dtype=dtype)
    tiles_x = (W + block_width - 1) // block_width
    tiles_y = (H + block_width - 1) // block_width
    ids_all = torch.from_numpy(np.ascontiguousarray(sorted_ids)).to(torch.int64)
    rows, cols = [], []
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            b, e = int(tile_bins[ty * tiles_x + tx, 0]), int(tile_bins[ty * tiles_x + tx, 1])
            if e <= b:
                continue
            ids = ids_all[b:e]
            if cls_filter >= 0:
                ids = ids[cls[ids] == cls_filter]
                if ids.numel() == 0:
                    continue
            y0, x0 = ty * block_width, tx * block_width
            y1, x1 = min(y0 + block_width, H), min(x0 + block_width, W)
            py = torch.arange(y0, y1, dtype=dtype) + 0.5
            px = torch.arange(x0, x1, dtype=dtype) + 0.5
            PY, PX = torch.meshgrid(py, px, indexing="ij")
            PX, PY = PX.reshape(-1, 1), PY.reshape(-1, 1)
            gx, gy = xys[ids, 0][None, :], xys[ids, 1][None, :]
            ca, cb, cc = conics[ids, 0][None, :], conics[ids, 1][None, :], conics[ids, 2][None, :]
            dx, dy = gx - PX, gy - PY
            sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
            raw = opac[ids][None, :] * torch.exp(-sigma)
            a = _ClampMaxST.apply(raw, alpha_clamp)
            valid = (sigma.detach() >= 0) & (a.detach() >= 1.0 / 255.0)
            a_eff = torch.where(valid, a, torch.zeros_like(a))
            one_minus = 1.0 - a_eff
            Tincl = torch.cumprod(one_minus, dim=1)
            Tex = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], dim=1)
            stop = valid & (Tincl.detach() <= 1e-4)
            done = torch.cummax(stop.to(torch.int8), dim=1)[0].bool()  # inclusive: the stopper is excluded
            contrib = valid & ~done
            wgt = torch.where(contrib, a * Tex, torch.zeros_like(a))
            pix = wgt @ colors[ids]
            Tfin = torch.prod(torch.where(contrib, one_minus, torch.ones_like(one_minus)), dim=1)
            img[y0:y1, x0:x1] = pix.reshape(y1 - y0, x1 - x0, C)
            alpha_out[y0:y1, x0:x1] = (1.0 - Tfin).reshape(y1 - y0, x1 - x0)
    return img, alpha_out


def render(frame, sorted_ids, tile_bins, sh_degree=3, sh_degree_to_use=None, block_width=16,
           alpha_clamp=0.999, dtype=torch.float64, class_renders=True, use_spec_exp=True):
    """Full forward on given tile lists.  Returns (leaves, dict) with raw outputs (differentiable)."""
    n = sh_degree if sh_degree_to_use is None else sh_degree_to_use
    leaves, cat = compose(frame, dtype)
    pr = project(cat, frame.camera, block_width, use_spec_exp=use_spec_exp)
    rgbs, opac = colours(cat, frame.camera, sh_degree, n)
    colors4 = torch.cat([rgbs, pr["depths"][:, None]], dim=1)
    img, alpha = blend(frame.camera, sorted_ids, tile_bins, pr["xys"], pr["conics"], colors4, opac,
                       block_width, alpha_clamp)
    out = dict(img=img, alpha=alpha, proj=pr, rgbs=rgbs, opac=opac, cat=cat)
    if class_renders:
        empty = torch.zeros(cat["means"].shape[0], 0, dtype=dtype)
        _, out["object_acc"] = blend(frame.camera, sorted_ids, tile_bins, pr["xys"], pr["conics"], empty, opac,
                                     block_width, alpha_clamp, cat["cls"], 1)
        _, out["background_acc"] = blend(frame.camera, sorted_ids, tile_bins, pr["xys"], pr["conics"], empty,
                                         opac, block_width, alpha_clamp, cat["cls"], 0)
    return leaves, out

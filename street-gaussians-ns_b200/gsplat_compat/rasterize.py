"""gsplat.rasterize.rasterize_gaussians (call sites street_gaussians_ns/sgn_splatfacto.py:954-967, :982-994).

N-channel colours are blended four channels per traversal (the fused main kernel carries rgb + depth,
i.e. four generic channels); binning happens once per call."""
from typing import Optional

import torch
from torch.autograd import Function

from .. import _lib, raster
from ._common import need_cuda


def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height: int, img_width: int,
                        block_width: int, background: Optional[torch.Tensor] = None, return_alpha: bool = False):
    """-> out_img[H,W,C] (, out_alpha[H,W]).  Differentiable w.r.t. xys, conics, colors, opacity."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    if colors.dtype == torch.uint8:
        colors = colors.float() / 255
    if background is not None:
        assert background.shape[0] == colors.shape[-1], f"incorrect shape of background color tensor, expected shape {colors.shape[-1]}"
    else:
        background = torch.ones(colors.shape[-1], dtype=torch.float32, device=colors.device)
    if xys.ndimension() != 2 or xys.size(1) != 2:
        raise ValueError("xys must have dimensions (N, 2)")
    if colors.ndimension() != 2:
        raise ValueError("colors must have dimensions (N, D)")
    out = _RasterizeGaussians.apply(xys.contiguous(), depths.contiguous(), radii.contiguous(), conics.contiguous(),
                                    num_tiles_hit.contiguous(), colors.contiguous(), opacity.contiguous(), img_height,
                                    img_width, block_width, background.contiguous())
    return (out[0], out[1]) if return_alpha else out[0]


def _camera(img_height, img_width, block_width) -> _lib.CameraStruct:
    cs = _lib.CameraStruct()
    cs.width, cs.height, cs.block_width = int(img_width), int(img_height), int(block_width)
    cs.sh_degree = cs.sh_degree_to_use = 0
    return cs


def _tile_bbox(xys, radii, cs):
    """gsplat get_tile_bbox on (xys, radii) as map_gaussian_to_intersects recomputes it (Appendix A.4)."""
    bw = float(cs.block_width)
    tiles_x = (cs.width + cs.block_width - 1) // cs.block_width
    tiles_y = (cs.height + cs.block_width - 1) // cs.block_width
    tc = xys / bw
    tr = (radii.float() / bw)[:, None]
    lim = torch.tensor([tiles_x, tiles_y], device=xys.device, dtype=torch.float32)
    tmin = torch.minimum(torch.trunc(tc - tr).clamp(min=0), lim)
    tmax = torch.minimum(torch.trunc(tc + tr + 1).clamp(min=0), lim)
    return torch.cat([tmin, tmax], dim=1).to(torch.int16).contiguous()


class _RasterizeGaussians(Function):
    @staticmethod
    def forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, block_width,
                background):
        L = _lib.load()
        xys, depths, conics = need_cuda(xys, "xys"), need_cuda(depths, "depths"), need_cuda(conics, "conics")
        colors, opacity, background = need_cuda(colors, "colors"), need_cuda(opacity, "opacity"), need_cuda(background, "background")
        N, Cc = colors.shape
        dev = xys.device
        cs = _camera(img_height, img_width, block_width)
        if block_width != 16:
            raise _lib.SgnError("the B200 blend kernels are specialised for block_width == 16 (the reference's setting)")
        radii = radii.to(torch.int32).contiguous()
        H, W = int(img_height), int(img_width)
        out_img = torch.empty(H, W, Cc, device=dev)
        out_alpha = torch.empty(H, W, device=dev)
        if int(num_tiles_hit.sum().item()) < 1:  # gsplat: no intersection -> background everywhere
            out_img[:] = background
            out_alpha.zero_()
            ctx.empty = True
            ctx.shapes = (xys.shape, conics.shape, colors.shape, opacity.shape)
            return out_img, out_alpha
        ctx.empty = False
        bbox = _tile_bbox(xys.detach(), radii, cs)
        groups = []
        bin_state = None
        opac = opacity.reshape(N)
        for c0 in range(0, Cc, 4):
            nch = min(4, Cc - c0)
            rec = torch.zeros(N, _lib.RECORD_FLOATS, device=dev)
            rec[:, 0:2], rec[:, 2:5], rec[:, 5] = xys.detach(), conics.detach(), opac.detach()
            rec[:, 6:6 + min(nch, 3)] = colors.detach()[:, c0:c0 + min(nch, 3)]
            rec[:, 9] = depths.detach()  # sort key
            if bin_state is None:
                bin_state = raster.bin_and_sort(cs, rec, radii, num_tiles_hit, bbox)
            M, sorted_ids, tile_bins = bin_state
            if nch == 4:
                rec[:, 9] = colors.detach()[:, c0 + 3]  # 4th channel rides in the depth slot (after binning)
            bo = _lib.BlendOpts()
            bo.alpha_clamp_fwd, bo.alpha_clamp_bwd = 0.999, 0.99
            bo.raw_mode = 1
            for k in range(4):
                bo.background[k] = float(background[c0 + k]) if k < nch else 0.0
            out = raster.blend_fwd(cs, bo, rec, sorted_ids, tile_bins, None)
            out_img[..., c0:c0 + min(nch, 3)] = out["rgb"][..., :min(nch, 3)]
            if nch == 4:
                out_img[..., c0 + 3] = out["depth"][..., 0]
            if c0 == 0:
                out_alpha.copy_(out["accumulation"][..., 0])
            groups.append((c0, nch, rec, bo, out))
        ctx.cs, ctx.groups, ctx.sorted_ids, ctx.tile_bins = cs, groups, sorted_ids, tile_bins
        ctx.N, ctx.Cc = N, Cc
        ctx.set_materialize_grads(False)
        return out_img, out_alpha

    @staticmethod
    def backward(ctx, v_out_img, v_out_alpha):
        if ctx.empty:
            z = [torch.zeros(s, device=v_out_img.device if v_out_img is not None else v_out_alpha.device) for s in ctx.shapes]
            return (z[0], None, None, z[1], None, z[2], z[3], None, None, None, None)
        N, Cc = ctx.N, ctx.Cc
        dev = ctx.sorted_ids.device
        v_xy = torch.zeros(N, 2, device=dev)
        v_conic = torch.zeros(N, 3, device=dev)
        v_colors = torch.zeros(N, Cc, device=dev)
        v_opac = torch.zeros(N, device=dev)
        for gi, (c0, nch, rec, bo, out) in enumerate(ctx.groups):
            v = dict(rgb=None, accumulation=None, depth=None, object_acc=None, background_acc=None)
            if v_out_img is not None:
                vr = torch.zeros(ctx.cs.height, ctx.cs.width, 3, device=dev)
                vr[..., :min(nch, 3)] = v_out_img[..., c0:c0 + min(nch, 3)]
                v["rgb"] = vr
                if nch == 4:
                    v["depth"] = v_out_img[..., c0 + 3].contiguous()
            if gi == 0 and v_out_alpha is not None:
                v["accumulation"] = v_out_alpha
            v_rec, _ = raster.blend_bwd(ctx.cs, bo, rec, ctx.sorted_ids, ctx.tile_bins, out, None, v, False)
            v_xy += v_rec[:, 0:2]
            v_conic += v_rec[:, 2:5]
            v_opac += v_rec[:, 5]
            v_colors[:, c0:c0 + min(nch, 3)] = v_rec[:, 6:6 + min(nch, 3)]
            if nch == 4:
                v_colors[:, c0 + 3] = v_rec[:, 9]
        return (v_xy, None, None, v_conic, None, v_colors, v_opac.reshape(-1, 1), None, None, None, None)

"""gsplat.sh (call site street_gaussians_ns/sgn_splatfacto.py:939, :292-293)."""

import torch
from torch.autograd import Function

from .. import _lib
from ._common import need_cuda, ptr, stream


def num_sh_bases(degree: int) -> int:
    if degree == 0:
        return 1
    if degree == 1:
        return 4
    if degree == 2:
        return 9
    if degree == 3:
        return 16
    return 25


def deg_from_sh(num_bases: int) -> int:
    return {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[num_bases]


def spherical_harmonics(degrees_to_use: int, viewdirs, coeffs, method: str = "poly"):
    """viewdirs[N,3] (unit), coeffs[N,K,3] -> colors[N,3]; differentiable w.r.t. coeffs only (gsplat 0.1.x)."""
    assert coeffs.shape[-2] >= num_sh_bases(degrees_to_use)
    assert degrees_to_use <= 3, "the B200 path implements SH up to degree 3 (the reference's sh_degree)"
    return _SphericalHarmonics.apply(degrees_to_use, viewdirs.contiguous(), coeffs.contiguous())


class _SphericalHarmonics(Function):
    @staticmethod
    def forward(ctx, degrees_to_use, viewdirs, coeffs):
        L = _lib.load()
        viewdirs, coeffs = need_cuda(viewdirs, "viewdirs"), need_cuda(coeffs, "coeffs")
        N, K = coeffs.shape[0], coeffs.shape[-2]
        if K > 16:
            raise ValueError("at most 16 SH coefficients (degree 3) are supported")
        colors = torch.empty(N, 3, device=coeffs.device)
        _lib.check(L.sgn_l1_sh(N, K, int(degrees_to_use), ptr(viewdirs), ptr(coeffs), None, ptr(colors), None, stream()),
                   "sgn_l1_sh")
        ctx.degree, ctx.K = int(degrees_to_use), K
        ctx.save_for_backward(viewdirs)
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        L = _lib.load()
        (viewdirs,) = ctx.saved_tensors
        N = viewdirs.shape[0]
        v_colors = v_colors.contiguous()
        v_coeffs = torch.empty(N, ctx.K, 3, device=v_colors.device)
        _lib.check(L.sgn_l1_sh(N, ctx.K, ctx.degree, ptr(viewdirs), None, ptr(v_colors), None, ptr(v_coeffs), stream()),
                   "sgn_l1_sh")
        return None, None, v_coeffs

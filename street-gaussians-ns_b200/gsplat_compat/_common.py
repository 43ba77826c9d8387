import ctypes as C

import numpy as np
import torch

from .. import _lib


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def need_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.SgnError(f"{name} must be a CUDA tensor: the B200 rasterizer has no CPU path")
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


def camera_struct(viewmat, fx, fy, cx, cy, img_height, img_width, block_width, clip_thresh=0.01) -> _lib.CameraStruct:
    cs = _lib.CameraStruct()
    vm = viewmat.detach().float().cpu().numpy().reshape(-1, 4)[:3].reshape(-1)
    for i in range(12):
        cs.viewmat[i] = float(vm[i])
    cs.fx, cs.fy, cs.cx, cs.cy = float(fx), float(fy), float(cx), float(cy)
    cs.width, cs.height = int(img_width), int(img_height)
    tan_x = np.float32(0.5 * int(img_width) / float(np.float32(fx)))
    tan_y = np.float32(0.5 * int(img_height) / float(np.float32(fy)))
    cs.limx, cs.limy = float(np.float32(1.3) * tan_x), float(np.float32(1.3) * tan_y)
    cs.clip_thresh = float(clip_thresh)
    cs.block_width = int(block_width)
    cs.sh_degree = 3
    cs.sh_degree_to_use = 3
    return cs

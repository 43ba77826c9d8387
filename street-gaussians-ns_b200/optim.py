"""Fused multi-tensor Adam over the flat gradient arena.

Drop-in for the nine ``torch.optim.Adam`` instances nerfstudio builds from
street_gaussians_ns/sgn_config.py:71-108 for the Gaussian parameter groups: same update rule (dense, eps 1e-15,
betas (0.9, 0.999)), one kernel launch for all ~200 tensors.  Moments live in two flat arenas that share
the layout of ``holder.grad_arena`` (raster.project_bwd), so a training step is
``forward_backward -> (all-reduce arena) -> FusedAdam.step(arena)``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib
from .scene import PARAM_NAMES

# street_gaussians_ns/sgn_config.py:84-107
REFERENCE_LRS: Dict[str, float] = {
    "means": 1.6e-4, "scales": 0.005, "quats": 0.001, "features_dc": 0.0025, "features_rest": 0.0025 / 20,
    "opacities": 0.05,
}

ADAM_DTYPE = np.dtype([("param", "<u8"), ("arena_offset", "<i8"), ("numel", "<i8"), ("chunk0", "<i4"),
                       ("beta1", "<f4"), ("beta2", "<f4"), ("eps", "<f4"), ("step_size", "<f4"),
                       ("sqrt_bc2", "<f4"), ("one_minus_beta1", "<f4"), ("one_minus_beta2", "<f4")])  # 56 B: the C struct is 8-byte aligned
assert ADAM_DTYPE.itemsize == C.sizeof(_lib.AdamTensor)


class FusedAdam:
    """``params``: per segment, the six parameter tensors in PARAM_NAMES order (the order of the gradient arena)."""

    def __init__(self, params: Sequence[Sequence[torch.Tensor]], lrs: Dict[str, float] = None, betas=(0.9, 0.999),
                 eps: float = 1e-15):
        self.L = _lib.load()
        self.lrs = dict(REFERENCE_LRS if lrs is None else lrs)
        self.betas, self.eps = betas, eps
        self.step_count = 0
        flat = [t for ps in params for t in ps]
        self.kinds = [PARAM_NAMES[i % 6] for i in range(len(flat))]
        self.device = flat[0].device
        sizes = [(t.numel() + 3) // 4 * 4 for t in flat]
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        chunk = self.L.sgn_adam_chunk_elems()
        chunks = np.array([(t.numel() + chunk - 1) // chunk for t in flat], np.int64)
        tab = np.zeros(len(flat), ADAM_DTYPE)
        tab["param"] = [t.data_ptr() for t in flat]
        tab["arena_offset"] = offsets
        tab["numel"] = [t.numel() for t in flat]
        tab["chunk0"] = np.concatenate([[0], np.cumsum(chunks)[:-1]])
        tab["beta1"], tab["beta2"], tab["eps"] = betas[0], betas[1], eps
        tab["one_minus_beta1"], tab["one_minus_beta2"] = 1.0 - betas[0], 1.0 - betas[1]  # double, then rounded (torch)
        self.table = tab
        self.num_chunks = int(chunks.sum())
        self.arena_elems = int(sum(sizes))
        self.exp_avg = torch.zeros(self.arena_elems, device=self.device)
        self.exp_avg_sq = torch.zeros(self.arena_elems, device=self.device)
        self._params = flat  # keep the tensors (and their storage) alive

    def set_lr(self, kind: str, lr: float) -> None:
        """Schedulers (e.g. the exponential decay of the means lr, sgn_config.py:85-90) update rates here."""
        self.lrs[kind] = lr

    def step(self, grad_arena: torch.Tensor) -> None:
        assert grad_arena.numel() >= self.arena_elems and grad_arena.is_cuda
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** self.step_count
        bc2 = 1.0 - b2 ** self.step_count
        tab = self.table
        tab["step_size"] = [self.lrs[k] / bc1 for k in self.kinds]
        tab["sqrt_bc2"] = math.sqrt(bc2)
        dev_tab = torch.from_numpy(tab.view(np.uint8).reshape(-1)).to(self.device, non_blocking=True)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.L.sgn_adam_step(C.c_void_p(dev_tab.data_ptr()), len(tab), self.num_chunks,
                                        C.c_void_p(grad_arena.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                                        C.c_void_p(self.exp_avg_sq.data_ptr()), stream), "sgn_adam_step")

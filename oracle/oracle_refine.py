"""ORACLE (test infrastructure, never a product path): torch restatement of the reference's refinement step for ONE
sub-model -- ``SplatfactoModel.refinement_after`` with ``cull_gaussians`` / ``split_gaussians`` / ``dup_gaussians`` and
the Adam-state surgery ``dup_in_optim`` / ``remove_from_optim``
(street_gaussians_ns/sgn_splatfacto.py:550-646, :648-672, :674-710, :712-720, :459-511).

Whole-tensor statements in the reference's ORDER (the order matters: ``split_gaussians`` rescales the split rows in
place before ``dups`` is evaluated, :696 then :582), on whatever device the tensors live on: the CPU tests run it on
CPU, the GPU test runs it on the GPU so that exp / log / sigmoid are the very CUDA functions the reference would call.

Parity status: PINNED.  This step is pure torch in the reference, so the reference's own code runs in the build
container: tests/golden/make_golden_reference.py executes ``SplatfactoModel.refinement_after`` itself (imported through
tests/golden/reference_loader.py) through every phase of the schedule with a live torch.optim.Adam state and commits
inputs, captured split samples and outputs as tests/golden/reference_vectors.npz; tests/test_reference_vectors.py holds
this restatement (and the product) to those vectors.  Hand-built known-answer cases are in tests/test_refine.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

PARAMS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")


@dataclass
class RefineConfig:
    """The fields of SplatfactoModelConfig the step reads (sgn_splatfacto.py:158-194 defaults; the scene graph's
    sub-model overrides are in sgn_config.py:49-65)."""

    warmup_length: int = 500
    refine_every: int = 100
    reset_alpha_every: int = 30
    stop_split_at: int = 15000
    stop_screen_size_at: int = 4000
    densify_grad_thresh: float = 0.0002
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    split_screen_size: float = 0.05
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    cull_screen_size: float = 0.15
    continue_cull_post_densification: bool = True


@dataclass
class SubModelState:
    params: Dict[str, torch.Tensor]                                   # the six gauss_params
    moments: Optional[Dict[str, Tuple[torch.Tensor, torch.Tensor]]]   # Adam exp_avg / exp_avg_sq per group, or None
    xys_grad_norm: Optional[torch.Tensor] = None
    vis_counts: Optional[torch.Tensor] = None
    max_2Dsize: Optional[torch.Tensor] = None


def quat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """gsplat 0.1.x ``_torch_impl.quat_to_rotmat`` (normalises, then the usual wxyz formula)."""
    w, x, y, z = torch.unbind(F.normalize(quat, dim=-1), dim=-1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, dim=-1).reshape(quat.shape[:-1] + (3, 3))


def _cull(st: SubModelState, cfg: RefineConfig, step: int, extra: Optional[torch.Tensor], record: dict) -> torch.Tensor:
    """cull_gaussians (:648-672): the mask of deleted rows; st.params are replaced by the survivors."""
    p = st.params
    culls = (torch.sigmoid(p["opacities"]) < cfg.cull_alpha_thresh).squeeze(-1)
    record["refine_culls_alpha_count"] = int(culls.sum())
    if extra is not None:
        culls = culls | extra
    if step > cfg.refine_every * cfg.reset_alpha_every:
        toobigs = torch.exp(p["scales"]).max(dim=-1).values > cfg.cull_scale_thresh
        if step < cfg.stop_screen_size_at:
            assert st.max_2Dsize is not None
            toobigs = toobigs | (st.max_2Dsize > cfg.cull_screen_size)
        culls = culls | toobigs
        record["refine_culls_toobigs_count"] = int(toobigs.sum())
    for name in PARAMS:
        p[name] = p[name][~culls]
    return culls


def _split(st: SubModelState, mask: torch.Tensor, samps: int, randn: Callable) -> Dict[str, torch.Tensor]:
    """split_gaussians (:674-710), including its in-place rescale of the split rows."""
    p = st.params
    n_splits = int(mask.sum())
    centered = randn(samps * n_splits)                                       # :680
    scaled = torch.exp(p["scales"][mask].repeat(samps, 1)) * centered        # :681-683
    quats = p["quats"][mask] / p["quats"][mask].norm(dim=-1, keepdim=True)   # :684
    rots = quat_to_rotmat(quats.repeat(samps, 1))                            # :685
    rotated = torch.bmm(rots, scaled[..., None]).squeeze(-1)                 # :686
    out = {"means": rotated + p["means"][mask].repeat(samps, 1)}             # :687
    out["features_dc"] = p["features_dc"][mask].repeat(samps, 1, 1)
    out["features_rest"] = p["features_rest"][mask].repeat(samps, 1, 1)
    out["opacities"] = p["opacities"][mask].repeat(samps, 1)
    size_fac = 1.6
    out["scales"] = torch.log(torch.exp(p["scales"][mask]) / size_fac).repeat(samps, 1)  # :695
    p["scales"][mask] = torch.log(torch.exp(p["scales"][mask]) / size_fac)               # :696 (in place)
    out["quats"] = p["quats"][mask].repeat(samps, 1)
    return out


def refinement_after(st: SubModelState, cfg: RefineConfig, step: int, last_size: Tuple[int, int], num_train_data: int,
                     randn: Optional[Callable] = None) -> dict:
    """One call of the reference's callback on one sub-model; mutates ``st``.  ``randn(k)`` must return ``[k,3]``
    standard-normal draws on the parameters' device (the reference calls ``torch.randn((k, 3), device=...)``)."""
    record: dict = {}
    if step <= cfg.warmup_length or st.xys_grad_norm is None:                 # :552-555
        return record
    p = st.params
    dev = p["means"].device
    if randn is None:
        randn = lambda k: torch.randn((k, 3), device=dev)  # noqa: E731
    reset_interval = cfg.reset_alpha_every * cfg.refine_every
    do_densification = step < cfg.stop_split_at and step % reset_interval > num_train_data + cfg.refine_every  # :563-566
    deleted = None
    if do_densification:
        avg_grad_norm = (st.xys_grad_norm / st.vis_counts) * 0.5 * max(last_size[0], last_size[1])  # :570
        high_grads = avg_grad_norm > cfg.densify_grad_thresh
        record["high_grads_count"] = int(high_grads.sum())
        splits = p["scales"].exp().max(dim=-1).values > cfg.densify_size_thresh
        if step < cfg.stop_screen_size_at:
            splits = splits | (st.max_2Dsize > cfg.split_screen_size)
        splits = splits & high_grads
        nsamps = cfg.n_split_samples
        split_params = _split(st, splits, nsamps, randn)
        record["refine_splits_count"] = int(splits.sum())
        dups = (p["scales"].exp().max(dim=-1).values <= cfg.densify_size_thresh) & high_grads  # after the in-place rescale
        dup_params = {name: p[name][dups] for name in PARAMS}
        record["refine_dups_count"] = int(dups.sum())
        for name in PARAMS:
            p[name] = torch.cat([p[name].detach(), split_params[name], dup_params[name]], dim=0)
        n_new = split_params["scales"].shape[0] + dup_params["scales"].shape[0]
        st.max_2Dsize = torch.cat([st.max_2Dsize, torch.zeros(n_new, device=dev)])
        if st.moments is not None:                                            # dup_in_optim x2 (:483-511)
            for name in PARAMS:
                m, v = st.moments[name]
                pad = (n_new,) + tuple(m.shape[1:])
                st.moments[name] = (torch.cat([m, torch.zeros(pad, device=dev)]), torch.cat([v, torch.zeros(pad, device=dev)]))
        splits_mask = torch.cat([splits, torch.zeros(n_new, dtype=torch.bool, device=dev)])  # :608-618
        deleted = _cull(st, cfg, step, splits_mask, record)
    elif step >= cfg.stop_split_at and cfg.continue_cull_post_densification:  # :620-621
        deleted = _cull(st, cfg, step, None, record)
    if deleted is not None and st.moments is not None:                        # remove_from_optim (:459-476)
        for name in PARAMS:
            m, v = st.moments[name]
            st.moments[name] = (m[~deleted], v[~deleted])
    if step < cfg.stop_split_at and step % reset_interval == cfg.refine_every:   # :629-642
        reset_value = cfg.cull_alpha_thresh * 2.0
        p["opacities"] = torch.clamp(p["opacities"], max=torch.logit(torch.tensor(reset_value, device=dev)).item())
        if st.moments is not None:
            m, v = st.moments["opacities"]
            st.moments["opacities"] = (torch.zeros_like(m), torch.zeros_like(v))
    st.xys_grad_norm = st.vis_counts = st.max_2Dsize = None                   # :644-646
    return record

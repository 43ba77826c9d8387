"""Generates tests/golden/refine_case.npz: known-answer vectors for the refinement step (split / duplicate / cull,
Adam moments carried along) of one small sub-model.

The reference cannot be imported in this container, so the vectors come from the torch restatement
of its statements (oracle/oracle_refine.py; street_gaussians_ns/sgn_splatfacto.py:550-720) run on the CPU.  The file
stores the INPUTS as well (parameters, moments, statistics, the standard-normal draws of the split samples), so it
does not depend on any random-number generator.  Run from the repo root:

    python tests/golden/make_golden_refine.py

CPU suite: the oracle still reproduces the file, and the product's row rules (g++ build) match it.  GPU suite: the
CUDA kernels match it.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_refine as orc  # noqa: E402

PARAMS = orc.PARAMS
STEP, SIZE, NTRAIN = 3400, (240, 320), 50  # densify + size culling + screen-size rules all active
CONFIG = dict(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refine_case.npz")


def make_inputs(n=400, F=5, seed=31):
    g = torch.Generator().manual_seed(seed)
    p = {"means": torch.randn(n, 3, generator=g) * 5, "scales": torch.randn(n, 3, generator=g) * 1.5 - 4.0,
         "quats": torch.randn(n, 4, generator=g), "features_dc": torch.randn(n, F, 3, generator=g),
         "features_rest": torch.randn(n, 15, 3, generator=g), "opacities": torch.randn(n, 1, generator=g) * 2.5 - 1.0}
    d = {"in_" + k: v.numpy() for k, v in p.items()}
    for k, v in p.items():
        d["in_m_" + k] = torch.randn(v.shape, generator=g).numpy()
        d["in_v_" + k] = torch.rand(v.shape, generator=g).numpy()
    vis = torch.randint(1, 9, (n,), generator=g).float()
    d["vis_counts"] = vis.numpy()
    d["xys_grad_norm"] = (torch.rand(n, generator=g) * vis * 2.5e-6).numpy()
    d["max_2Dsize"] = (torch.rand(n, generator=g) * 0.2).numpy()
    d["samples"] = torch.randn(2 * n, 3, generator=g).numpy()  # more than any split needs; the first 2*n_splits rows are used
    return d


def run_oracle(d):
    """Returns the oracle's outputs for the inputs ``d`` (dict of numpy arrays)."""
    st = orc.SubModelState({k: torch.from_numpy(d["in_" + k].copy()) for k in PARAMS},
                           {k: (torch.from_numpy(d["in_m_" + k].copy()), torch.from_numpy(d["in_v_" + k].copy())) for k in PARAMS},
                           torch.from_numpy(d["xys_grad_norm"].copy()), torch.from_numpy(d["vis_counts"].copy()),
                           torch.from_numpy(d["max_2Dsize"].copy()))
    samples = torch.from_numpy(d["samples"])
    rec = orc.refinement_after(st, orc.RefineConfig(**CONFIG), STEP, SIZE, NTRAIN, randn=lambda k: samples[:k].clone())
    out = {"out_" + k: st.params[k].numpy() for k in PARAMS}
    for k in PARAMS:
        out["out_m_" + k], out["out_v_" + k] = st.moments[k][0].numpy(), st.moments[k][1].numpy()
    out["counts"] = np.array([rec["high_grads_count"], rec["refine_splits_count"], rec["refine_dups_count"],
                              rec["refine_culls_alpha_count"], rec["refine_culls_toobigs_count"]], np.int64)
    return out


if __name__ == "__main__":
    d = make_inputs()
    d.update(run_oracle(d))
    np.savez_compressed(OUT, **d)
    print(OUT, {k: int(v) for k, v in zip(("high", "splits", "dups", "alpha", "toobig"), d["counts"])},
          "rows", d["in_means"].shape[0], "->", d["out_means"].shape[0])

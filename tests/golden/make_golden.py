"""Generates tests/golden/tiny_scene.npz: known-answer vectors for the hot path on a tiny seeded scene.

The reference (street-gaussians-ns + gsplat 0.1.x) cannot be imported in this container, so
these vectors come from the oracle restatement: forward images from the C oracle (float32, gsplat arithmetic),
parameter gradients from the float64 torch/autograd oracle in the consistent-clamp mode AND from the C oracle
in gsplat's clamp mode (forward 0.999 / backward 0.99).  Run from the repo root:

    python tests/golden/make_golden.py

The CPU suite checks the oracle still reproduces the file (regression pin); the GPU suite checks the CUDA path
against it without needing the oracle's build.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from oracle import oracle_c, oracle_torch  # noqa: E402

SCENE = dict(n_background=1500, n_actors=1, n_per_actor=300, width=64, height=48, seed=12,
             actor_shift=np.array([1.75, 0.4, 2.0]))


def build():
    fr = syn.make_frame(**SCENE)
    orc = oracle_c.Oracle(fr)  # gsplat clamps
    fw = orc.forward()
    H, W = fr.camera.height, fr.camera.width
    g = torch.Generator().manual_seed(21)
    # cotangents only where no decision is fragile and the clamp(rgb, max=1) post-op is inactive, and none on the
    # depth channel: then d(loss)/d(raw outputs) == d(loss)/d(model outputs) and the file pins both levels
    ok = ((fw.fragile == 0) & (fw.fragile_obj == 0) & (fw.fragile_bg == 0) & (fw.img[..., :3].max(-1) < 0.999)).astype(np.float32)
    w_img = (torch.rand(H, W, 4, generator=g).numpy() * ok[..., None]).astype(np.float32)
    w_img[..., 3] = 0.0
    w_a = (torch.rand(H, W, generator=g).numpy() * ok).astype(np.float32)
    w_o = (torch.rand(H, W, generator=g).numpy() * ok).astype(np.float32)
    grads, _ = orc.backward(fw, w_img, w_a, w_o, None)
    out = dict(
        img=fw.img, alpha=(1 - fw.final_T), object_acc=(1 - fw.obj_T), background_acc=(1 - fw.bg_T),
        fragile=(fw.fragile | fw.fragile_obj | fw.fragile_bg), radii=fw.radii, num_tiles_hit=fw.num_tiles_hit,
        xys=fw.xys, conics=fw.conics, depths=fw.depths, w_img=w_img, w_a=w_a, w_o=w_o,
        sorted_ids=fw.sorted_ids, tile_bins=fw.tile_bins)
    for si, gr in enumerate(grads):
        for k, v in gr.items():
            out[f"grad_{si}_{k}"] = v
    return out


if __name__ == "__main__":
    out = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_scene.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; M =", len(out["sorted_ids"]))

"""Generates tests/golden/reference_glue_tiny.npz: what THE REFERENCE'S OWN ``SplatfactoSceneGraphModel.get_outputs`` and its
autograd backward produce on a tiny seeded scene when gsplat's three calls are served by the oracle's restatement
(tests/golden/reference_glue.py explains the construction).  Three runs:

  train   training mode, no sky: rgb / accumulation / depth / object_acc / background_acc, the side-effect attributes
          (xys, depths, radii, conics, num_tiles_hit per model), and the parameter gradients of the fixed linear loss
          sum(w*rgb) + sum(v*accumulation) + sum(u*object_acc) for every sub-model;
  eval    eval mode with a sky image: sky blend, clamp(0,1), background_rgb / object_rgb;
  sched   training at step 1500 (SH degree 1 of 3 by the schedule), rotated camera of the rig, 70x50 image (ragged tiles);
  empty   camera turned away from everything: the base model's early-out dict (sgn_splatfacto.py:878-886); the scene-graph
          wrapper itself raises in this situation (recorded).

    python tests/golden/make_golden_reference_glue.py        (only where /root/reference is mounted)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_glue as rg  # noqa: E402  (puts the repo root on sys.path)
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402

OUT = os.path.join(HERE, "reference_glue_tiny.npz")
SCENE = dict(n_background=1500, n_actors=2, n_per_actor=300, width=64, height=48, seed=12, actor_shift=np.array([1.75, 0.4, 2.0]))
PARAMS = rg.PARAMS
OUTPUTS = ("rgb", "accumulation", "depth", "object_acc", "background_acc")


def scene():
    return syn.make_frame(**SCENE)


SCHED_STEP = 1500  # training: n = min(step // sh_degree_interval, sh_degree) = 1 (sgn_splatfacto.py:936-938)
SCENE2 = dict(n_background=2500, n_actors=3, n_per_actor=250, width=70, height=50, seed=21, actor_shift=np.array([0.5, 0.0, -2.0]))


def scene2():
    return syn.make_frame(c2w=syn.waymo_rig(4)[6], **SCENE2)


def cotangents(H, W):
    g = torch.Generator().manual_seed(21)
    return torch.rand(H, W, 3, generator=g), torch.rand(H, W, 1, generator=g), torch.rand(H, W, 1, generator=g)


def sky_image(H, W):
    return torch.rand(H, W, 3, generator=torch.Generator().manual_seed(4))


def away_camera():
    """Looks along +z (the scene lies along -z): nothing projects in front of the camera."""
    c2w = np.concatenate([np.diag([-1.0, 1.0, -1.0]), np.zeros((3, 1))], axis=1)
    return syn.make_camera(SCENE["width"], SCENE["height"], c2w=c2w, time=21.0)


def build():
    d = {}
    fr = scene()
    H, W = fr.camera.height, fr.camera.width
    # ---- train -------------------------------------------------------------------------------------------------
    m, cam = rg.build_reference_model(fr, training=True)
    out = m.get_outputs(cam)
    w, v, u = cotangents(H, W)
    loss = (out["rgb"] * w).sum() + (out["accumulation"] * v).sum() + (out["object_acc"] * u).sum()
    loss.backward()
    for k in OUTPUTS:
        d["train_" + k] = out[k].detach().numpy()
    for k in ("xys", "depths", "radii", "conics", "num_tiles_hit"):
        d["train_side_" + k] = getattr(m, k).detach().numpy()
    d["train_xys_grad"] = torch.cat([m.all_models[n].xys.grad for n in m.visible_model_names]).numpy()
    for si, name in enumerate(m.visible_model_names):
        sub = m.all_models[name]
        assert sub.xys.shape[0] == sub.num_points and sub.last_size == (H, W)
        for k in PARAMS:
            d[f"train_grad_{si}_{k}"] = sub.gauss_params[k].grad.numpy()
    d["train_visible"] = np.array(m.visible_model_names)
    # ---- eval with sky --------------------------------------------------------------------------------------------
    sky = sky_image(H, W)
    m, cam = rg.build_reference_model(fr, training=False, sky=sky)
    with torch.no_grad():
        out = m.get_outputs(cam)
    for k in OUTPUTS + ("sky", "background_rgb", "object_rgb"):
        d["eval_" + k] = out[k].numpy()
    # ---- SH-degree schedule, rotated rig camera, ragged image size -------------------------------------------------------
    fr3 = scene2()
    m, cam = rg.build_reference_model(fr3, training=True, step=SCHED_STEP)
    out = m.get_outputs(cam)
    for k in OUTPUTS:
        d["sched_" + k] = out[k].detach().numpy()
    d["sched_radii"] = m.radii.numpy()
    # ---- nothing visible ------------------------------------------------------------------------------------------
    fr2 = scene()
    fr2.camera = away_camera()
    m, cam = rg.build_reference_model(fr2, training=False)
    # The base model returns its early-out dict (sgn_splatfacto.py:878-886), but the scene graph then goes on to render the
    # sub-model accumulations and trips ``assert (self.num_tiles_hit > 0).any()`` (:944): with nothing in view the
    # reference raises.  Recorded as such; the base model's early-out is captured by calling it directly.
    try:
        with torch.no_grad():
            m.get_outputs(cam)
        d["empty_scene_graph_raises"] = np.array(False)
    except AssertionError:
        d["empty_scene_graph_raises"] = np.array(True)
    with torch.no_grad():
        out = type(m).__mro__[1].get_outputs(m, cam)   # SplatfactoModel.get_outputs on the composed tensors
    d["empty_keys"] = np.array(sorted(out.keys()))
    for k in ("rgb", "accumulation", "depth"):
        d["empty_" + k] = out[k].numpy()
    return d


if __name__ == "__main__":
    d = build()
    np.savez_compressed(OUT, **d)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB;", "empty-scene keys:", list(d["empty_keys"]))

"""Generates tests/golden/reference_vectors.npz by EXECUTING THE REFERENCE'S OWN CODE on the CPU
(/root/reference/street_gaussians_ns/sgn_splatfacto.py and sgn_splatfacto_scene_graph.py, imported through
tests/golden/reference_loader.py -- read its header for what is stubbed and why):

  refine_*   ``SplatfactoModel.refinement_after`` (sgn_splatfacto.py:550-646, with cull / split / dup and the Adam-state
             surgery :459-511) on one seeded sub-model with a live per-group ``torch.optim.Adam`` state, through every
             phase of the schedule; the standard-normal draws of ``split_gaussians`` are captured and stored;
  loss_*     ``SplatfactoSceneGraphModel.get_loss_dict`` (:1042-1094 + scene graph :376-391): L1 (with and without
             mask), sky accumulation, object-accumulation entropy (the SSIM term is not in the fixture);
  idft_*     ``IDFT`` and ``get_fourier_features`` (scene graph :420-433, :239-247);
  stats_*    ``after_train`` (:513-541): the running densification statistics over two steps;
  view_*     what ``get_outputs`` (:793-873) hands to gsplat's ``project_gaussians``: world->camera matrix, intrinsics,
             image size, block width, ``exp(scales)``, unit quaternions (the call itself is intercepted).

This is the one part of the path where the reference itself -- not a restatement -- can run in the build container
(pure torch, no gsplat / nerfstudio arithmetic), so these vectors PIN the refinement oracle, the product's row rules,
the loss expressions and the Fourier basis against the reference.  Only runs where /root/reference is mounted:

    python tests/golden/make_golden_reference.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # repo root: the seeded synthetic camera rig
import reference_loader as rl  # noqa: E402

OUT = os.path.join(HERE, "reference_vectors.npz")
PARAMS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")
SIZE, NTRAIN = (240, 320), 50
CONFIG = dict(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2)  # the scene graph's background sub-model (sgn_config.py:46-56)
# (label, step, config overrides)
REFINE_CASES = [("densify_screen", 700, {}), ("densify_screen_big", 3400, {}), ("densify_3samples", 7400, {"n_split_samples": 3}),
                ("reset_only", 3100, {}), ("cull_only", 25000, {})]


def refine_inputs(n=96, F=5, seed=77):
    g = torch.Generator().manual_seed(seed)
    d = {"means": torch.randn(n, 3, generator=g) * 5, "scales": torch.randn(n, 3, generator=g) * 1.5 - 4.0,
         "quats": torch.randn(n, 4, generator=g), "features_dc": torch.randn(n, F, 3, generator=g),
         "features_rest": torch.randn(n, 15, 3, generator=g), "opacities": torch.randn(n, 1, generator=g) * 2.5 - 1.0}
    out = {"refine_in_" + k: v.numpy() for k, v in d.items()}
    for k, v in d.items():
        out["refine_in_m_" + k] = torch.randn(v.shape, generator=g).numpy()
        out["refine_in_v_" + k] = torch.rand(v.shape, generator=g).numpy()
    vis = torch.randint(1, 9, (n,), generator=g).float()
    out["refine_vis_counts"] = vis.numpy()
    out["refine_xys_grad_norm"] = (torch.rand(n, generator=g) * vis * 2.5e-6).numpy()
    out["refine_max_2Dsize"] = (torch.rand(n, generator=g) * 0.2).numpy()
    return out


def run_reference_refinement(base, inp, step, overrides, seed=123):
    cfg = base.SplatfactoModelConfig(**{**CONFIG, **overrides})
    m = rl.bare_model(base, {k: torch.from_numpy(inp["refine_in_" + k].copy()) for k in PARAMS}, cfg, step=step, num_train_data=NTRAIN)
    groups = {}
    for k in PARAMS:  # one Adam per group, this sub-model at index 0 of its parameter list (nerfstudio's layout)
        p = m.gauss_params[k]
        opt = torch.optim.Adam([p], lr=1e-3, eps=1e-15)
        opt.state[p] = {"step": torch.tensor(7.0), "exp_avg": torch.from_numpy(inp["refine_in_m_" + k].copy()),
                        "exp_avg_sq": torch.from_numpy(inp["refine_in_v_" + k].copy())}
        groups[k] = opt
    m.xys_grad_norm = torch.from_numpy(inp["refine_xys_grad_norm"].copy())
    m.vis_counts = torch.from_numpy(inp["refine_vis_counts"].copy())
    m.max_2Dsize = torch.from_numpy(inp["refine_max_2Dsize"].copy())
    m.last_size = SIZE
    drawn = []
    real_randn = torch.randn

    def recording_randn(*a, **k):
        t = real_randn(*a, **k)
        drawn.append(t.clone())
        return t

    torch.manual_seed(seed)
    torch.randn = recording_randn
    try:
        m.refinement_after(rl.Optimizers(groups), step)
    finally:
        torch.randn = real_randn
    out = {k: m.gauss_params[k].detach().numpy().copy() for k in PARAMS}
    for k in PARAMS:
        st = groups[k].state[groups[k].param_groups[0]["params"][0]]
        out["m_" + k], out["v_" + k] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
        assert float(st["step"]) == 7.0
    assert len(drawn) <= 1
    out["samples"] = drawn[0].numpy() if drawn else np.zeros((0, 3), np.float32)
    rec = m.refine_record_dict
    out["counts"] = np.array([rec.get("high_grads_count", -1), rec.get("refine_splits_count", -1), rec.get("refine_dups_count", -1),
                              rec.get("refine_culls_alpha_count", -1), rec.get("refine_culls_toobigs_count", -1)], np.int64)
    assert m.xys_grad_norm is None and m.vis_counts is None and m.max_2Dsize is None
    return out


def loss_vectors(base, graph):
    g = torch.Generator().manual_seed(5)
    H, W = 24, 40
    d = {"loss_rgb": torch.rand(H, W, 3, generator=g), "loss_gt": torch.rand(H, W, 3, generator=g),
         "loss_mask": (torch.rand(H, W, 1, generator=g) > 0.3).float(), "loss_accumulation": torch.rand(H, W, 1, generator=g),
         "loss_object_acc": torch.rand(H, W, 1, generator=g), "loss_semantic": torch.randint(0, 3, (H, W, 1), generator=g)}
    d["loss_object_acc"][0, :5] = 0.0   # exercises the clamp to [1e-5, 1 - 1e-5]
    d["loss_object_acc"][1, :5] = 1.0
    m = graph.SplatfactoSceneGraphModel.__new__(graph.SplatfactoSceneGraphModel)
    torch.nn.Module.__init__(m)
    m.config = graph.SplatfactoSceneGraphModelConfig()
    m.ssim = sys.modules["pytorch_msssim"].SSIM()
    m.step = m.config.background_model.stop_split_at + 1   # the entropy term is live (scene graph :386)
    m.train()
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self      # `sky_mask = (...).cuda()` (sgn_splatfacto.py:1091) on a CPU-only box
    try:
        out = {}
        for tag, with_mask in (("plain", False), ("masked", True)):
            batch = {"image": d["loss_gt"].clone(), "semantic": d["loss_semantic"].clone()}
            if with_mask:
                batch["mask"] = d["loss_mask"].clone()
            outputs = {"rgb": d["loss_rgb"].clone(), "accumulation": d["loss_accumulation"].clone(),
                       "object_acc": d["loss_object_acc"].clone()}
            losses = m.get_loss_dict(outputs, batch)
            out[f"loss_{tag}"] = np.array([float(losses["Ll1"]), float(losses["sky_accumulation"]),
                                           float(losses["object_acc_entropy_loss"])], np.float64)
    finally:
        torch.Tensor.cuda = real_cuda
    out["loss_weights"] = np.array([m.config.ssim_lambda, m.config.sky_acc_loss_mult, m.config.object_acc_entropy_loss_mult,
                                    m.config.background_model.stop_split_at], np.float64)
    out.update({k: v.numpy() for k, v in d.items()})
    return out


def fourier_vectors(base, graph):
    g = torch.Generator().manual_seed(9)
    dc = torch.randn(50, 5, 3, generator=g)
    cases = [(21, list(range(85)), 5, 1.0), (0, [3, 4, 5], 5, 1.0), (7, [7], 5, 1.0), (40, list(range(10, 60)), 3, 1.0),
             (12, list(range(85)), 5, 2.0), (5, list(range(20)), 1, 1.0)]
    out = {"idft_features_dc": dc.numpy(), "idft_cases": np.array([(f, fl[0], fl[-1], len(fl), dim, sc) for f, fl, dim, sc in cases], np.float64)}
    for i, (frame, frame_list, dim, scale) in enumerate(cases):
        obj = types.SimpleNamespace(config=types.SimpleNamespace(fourier_features_scale=scale, fourier_features_dim=dim),
                                    features_dc=dc[:, :dim])
        me = types.SimpleNamespace(object_annos=types.SimpleNamespace(objects_frames={"t": frame_list}), device=torch.device("cpu"))
        feat = graph.SplatfactoSceneGraphModel.get_fourier_features(me, frame, "t", obj)
        if len(frame_list) == 1:
            t = 1.0 * scale
        else:
            t = (frame - frame_list[0]) / (frame_list[-1] - frame_list[0]) * scale
        out[f"idft_basis_{i}"] = graph.IDFT(t, dim).numpy()[0]
        out[f"idft_feat_{i}"] = feat.numpy()
    return out


def after_train_vectors(base):
    """``SplatfactoModel.after_train`` (sgn_splatfacto.py:513-541) twice on one sub-model: the first call creates the
    running statistics, the second accumulates on the rows that were visible."""
    g = torch.Generator().manual_seed(14)
    n = 300
    out = {"stats_size": np.array([240, 320])}
    m = rl.bare_model(base, {"means": torch.zeros(n, 3), "scales": torch.zeros(n, 3), "quats": torch.ones(n, 4),
                             "features_dc": torch.zeros(n, 1, 3), "features_rest": torch.zeros(n, 15, 3), "opacities": torch.zeros(n, 1)},
                      step=1000)
    m.last_size = (240, 320)
    for call in range(2):
        radii = (torch.rand(n, generator=g) * 40).to(torch.int32) * (torch.rand(n, generator=g) > 0.35).to(torch.int32)
        grad = torch.randn(n, 2, generator=g) * 1e-4
        m.xys = torch.zeros(n, 2, requires_grad=True)
        m.xys.grad = grad.clone()
        m.radii = radii.clone()
        m.after_train(1000)
        out[f"stats_radii_{call}"], out[f"stats_xys_grad_{call}"] = radii.numpy(), grad.numpy()
        out[f"stats_xys_grad_norm_{call}"] = m.xys_grad_norm.numpy().copy()
        out[f"stats_vis_counts_{call}"] = m.vis_counts.numpy().copy()
        out[f"stats_max_2Dsize_{call}"] = m.max_2Dsize.numpy().copy()
    return out


class _Captured(Exception):
    pass


def view_vectors(base):
    """Runs the reference's ``get_outputs`` (sgn_splatfacto.py:793-873) up to its ``project_gaussians`` call and captures
    the arguments it passes: the world->camera matrix built from an OpenGL camera_to_worlds (:822-836), the intrinsics,
    image size, block width, glob_scale, and the pre-ops ``exp(scales)`` / ``quats / ||quats||`` (:857-864)."""
    import street_gaussians_ns_b200.synthetic as syn
    Cameras = sys.modules["nerfstudio.cameras.cameras"].Cameras
    rig = syn.waymo_rig(4)
    g = torch.Generator().manual_seed(3)
    rand_rot, _ = np.linalg.qr(torch.randn(3, 3, generator=g).numpy().astype(np.float64))
    poses = [np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1), rig[7], rig[13],
             np.concatenate([rand_rot, np.array([[1.5], [-0.7], [3.25]])], axis=1)]
    n = 40
    params = {"means": torch.randn(n, 3, generator=g), "scales": torch.randn(n, 3, generator=g) - 3, "quats": torch.randn(n, 4, generator=g),
              "features_dc": torch.randn(n, 1, 3, generator=g), "features_rest": torch.randn(n, 15, 3, generator=g),
              "opacities": torch.randn(n, 1, generator=g)}
    out = {"view_in_" + k: v.numpy() for k, v in params.items()}
    captured = {}

    def capture(means, scales, glob_scale, quats, viewmat, fx, fy, cx, cy, H, W, block_width, *rest):
        captured.update(means=means, scales=scales, glob_scale=glob_scale, quats=quats, viewmat=viewmat, fx=fx, fy=fy, cx=cx, cy=cy,
                        H=H, W=W, block_width=block_width)
        raise _Captured()

    m = rl.bare_model(base, params)
    m.back_color = torch.zeros(3)
    m.crop_box = None
    m.train()
    real = base.project_gaussians
    base.project_gaussians = capture
    try:
        for i, c2w in enumerate(poses):
            W, H = (1920, 1280) if i % 2 == 0 else (640, 480)
            f = np.float32(2055.0 * W / 1920.0)
            cam = Cameras()
            cam.shape = (1,)
            cam.camera_to_worlds = torch.from_numpy(c2w.astype(np.float32))[None]
            cam.fx, cam.fy = torch.tensor([[f]]), torch.tensor([[f * np.float32(1.01)]])
            cam.cx, cam.cy = torch.tensor([[W / 2.0 + 0.25]]), torch.tensor([[H / 2.0 - 0.5]])
            cam.width, cam.height = torch.tensor([[W]]), torch.tensor([[H]])
            cam.rescale_output_resolution = lambda s: None
            try:
                m.get_outputs(cam)
                raise AssertionError("the reference did not reach project_gaussians")
            except _Captured:
                pass
            out[f"view_c2w_{i}"] = c2w.astype(np.float32)
            out[f"view_viewmat_{i}"] = captured["viewmat"].detach().numpy().copy()
            out[f"view_scalars_{i}"] = np.array([captured["fx"], captured["fy"], captured["cx"], captured["cy"], captured["H"], captured["W"],
                                                 captured["block_width"], captured["glob_scale"]], np.float64)
            assert m.last_size == (H, W)
    finally:
        base.project_gaussians = real
    out["view_num_cameras"] = np.array(len(poses))
    out["view_exp_scales"] = captured["scales"].detach().numpy().copy()
    out["view_unit_quats"] = captured["quats"].detach().numpy().copy()
    assert torch.equal(captured["means"], m.gauss_params["means"])
    return out


def build():
    base, graph = rl.load()
    d = refine_inputs()
    for label, step, over in REFINE_CASES:
        res = run_reference_refinement(base, d, step, over)
        d.update({f"refine_{label}_{k}": v for k, v in res.items()})
    d.update(loss_vectors(base, graph))
    d.update(fourier_vectors(base, graph))
    d.update(view_vectors(base))
    d.update(after_train_vectors(base))
    return d


if __name__ == "__main__":
    d = build()
    np.savez_compressed(OUT, **d)
    for label, _, _ in REFINE_CASES:
        print(label, "rows", d["refine_in_means"].shape[0], "->", d[f"refine_{label}_means"].shape[0], "counts", d[f"refine_{label}_counts"].tolist())
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")

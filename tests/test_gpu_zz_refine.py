"""Refinement kernels (csrc/refine.cu: sgn_refine_decide / sgn_refine_apply) on the B200 against
the torch restatement of the reference's ``refinement_after`` (oracle/oracle_refine.py; street_gaussians_ns/
sgn_splatfacto.py:550-720) run ON THE SAME GPU, i.e. with the very CUDA exp / log / sigmoid / division kernels the
reference would execute.  Structure (which rows survive / split / duplicate, row order), copied values and Adam moments
are compared exactly; split-sample means (the reference rotates with torch.bmm) and shrunk scales to fp32 rounding.

Also here: FusedAdam over a gradient arena that only holds the visible sub-models, against torch.optim.Adam with
``grad = None`` for the absent ones; and a training step after a refinement (new row counts through the whole path).

(The file sorts last on purpose: these kernels were written in a GPU-less container and verified through a g++ build of
their row rules, tests/test_refine.py.)"""
import numpy as np
import pytest
import torch

import street_gaussians_ns_b200.synthetic as syn
from oracle import oracle_refine as orc
from street_gaussians_ns_b200 import refine
from street_gaussians_ns_b200.model import ActorPose, SceneGraphConfig, SceneGraphRasterModel
from street_gaussians_ns_b200.optim import REFERENCE_LRS, FusedAdam
from street_gaussians_ns_b200.scene import PARAM_NAMES

pytestmark = pytest.mark.gpu


def make_state(n, F, seed, dev):
    g = torch.Generator().manual_seed(seed)
    p = {"means": torch.randn(n, 3, generator=g) * 5, "scales": torch.randn(n, 3, generator=g) * 1.5 - 4.0,
         "quats": torch.randn(n, 4, generator=g), "features_dc": torch.randn(n, F, 3, generator=g),
         "features_rest": torch.randn(n, 15, 3, generator=g), "opacities": torch.randn(n, 1, generator=g) * 2.5 - 1.0}
    m = {k: (torch.randn(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in p.items()}
    vis = torch.randint(1, 9, (n,), generator=g).float()
    xgn = torch.rand(n, generator=g) * vis * 2.5e-6
    m2d = torch.rand(n, generator=g) * 0.2
    return orc.SubModelState({k: v.to(dev) for k, v in p.items()}, {k: (a.to(dev), b.to(dev)) for k, (a, b) in m.items()},
                             xgn.to(dev), vis.to(dev), m2d.to(dev))


def clone_state(st):
    return orc.SubModelState({k: v.clone() for k, v in st.params.items()}, {k: (a.clone(), b.clone()) for k, (a, b) in st.moments.items()},
                             st.xys_grad_norm.clone(), st.vis_counts.clone(), st.max_2Dsize.clone())


def assert_same(new_p, new_m, ost, label):
    for k, name in enumerate(PARAM_NAMES):
        a, b = new_p[k], ost.params[name]
        assert a.shape == b.shape, (label, name, tuple(a.shape), tuple(b.shape))
        if name in ("means", "scales"):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5, msg=lambda m: f"{label} {name}: {m}")
        else:
            assert torch.equal(a, b), (label, name)
        if new_m is not None and ost.moments is not None:
            assert torch.equal(new_m[k][0], ost.moments[name][0]), (label, name, "exp_avg")
            assert torch.equal(new_m[k][1], ost.moments[name][1]), (label, name, "exp_avg_sq")


PHASES = [("densify+screen", 700, {}), ("densify+screen+big", 3400, {}), ("densify", 4400, {}),
          ("densify+3samples", 7400, {"n_split_samples": 3}), ("reset-only", 3100, {}), ("cull-only", 25000, {})]


@pytest.mark.parametrize("label,step,over", PHASES, ids=[p[0] for p in PHASES])
@pytest.mark.parametrize("F", [1, 5])
def test_refine_kernels_match_reference_statements(label, step, over, F):
    dev = torch.device("cuda", 0)
    cfg = orc.RefineConfig(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2, **over)
    st = make_state(20011, F, 5 + F, dev)
    size, ntrain = (240, 320), 50
    mine = clone_state(st)
    settings = refine.RefineSettings(**{k: getattr(cfg, k) for k in refine.RefineSettings.__dataclass_fields__})
    g = torch.Generator(device=dev).manual_seed(17)
    new_p, new_m, plan = refine.refine_tensors([mine.params[k] for k in PARAM_NAMES], [mine.moments[k] for k in PARAM_NAMES],
                                               (mine.xys_grad_norm, mine.vis_counts, mine.max_2Dsize), settings, step, size, ntrain,
                                               generator=g)
    og = torch.Generator(device=dev).manual_seed(17)
    rec = orc.refinement_after(st, cfg, step, size, ntrain, randn=lambda k: torch.randn((k, 3), device=dev, generator=og))
    torch.cuda.synchronize()
    assert_same(new_p, new_m, st, f"{label} F={F}")
    if label.startswith("densify"):
        assert plan.totals[3] > 500 and plan.totals[2] > 500 and plan.totals[0] < plan.n
        got = plan.record()
        for k in ("high_grads_count", "refine_splits_count", "refine_dups_count", "refine_culls_alpha_count"):
            assert got[k] == rec[k], (k, got, rec)
        # the untouched inputs: apply reads, never writes, its sources
        assert mine.params["means"].shape[0] == 20011


def test_refine_argument_errors():
    from street_gaussians_ns_b200 import _lib
    dev = torch.device("cuda", 0)
    cfg = refine.make_config(refine.RefineSettings(), 700, (64, 48), True)
    cfg.n_split_samples = 0
    z = torch.zeros(8, 3, device=dev)
    with pytest.raises(_lib.SgnError, match="n_split_samples"):
        refine.plan_submodel(z, torch.zeros(8, 1, device=dev), torch.zeros(8, device=dev), torch.ones(8, device=dev),
                             torch.zeros(8, device=dev), cfg)


def test_fused_adam_skips_absent_submodels_like_torch():
    """Actor 1 has no gradient in steps 1 and 3 (not in view): torch.optim.Adam leaves its moments and step count alone."""
    fr = syn.make_frame(n_background=3001, n_actors=2, n_per_actor=503, width=64, height=48, seed=6)
    cpu_params = [[t.detach().clone() for t in s.params.tensors()] for s in fr.segments]
    gpu_params = [[t.detach().clone().cuda() for t in s.params.tensors()] for s in fr.segments]
    opt = FusedAdam(gpu_params)
    ref_opts = {k: torch.optim.Adam([ps[i] for ps in cpu_params], lr=REFERENCE_LRS[k], eps=1e-15) for i, k in enumerate(PARAM_NAMES)}
    g = torch.Generator().manual_seed(0)
    for step in range(5):
        present = [0, 1, 2] if step % 2 == 0 else [0, 2]
        chunks = []
        for si, ps in enumerate(cpu_params):
            for t in ps:
                if si in present:
                    gr = torch.randn(t.shape, generator=g)
                    t.grad = gr.clone()
                    pad = (-gr.numel()) % 4
                    chunks.append(torch.cat([gr.reshape(-1), torch.zeros(pad)]))
                else:
                    t.grad = None
        for o in ref_opts.values():
            o.step()
        opt.step(torch.cat(chunks).cuda(), present=None if len(present) == 3 else present)
    torch.cuda.synchronize()
    for ps_c, ps_g in zip(cpu_params, gpu_params):
        for name, a, b in zip(PARAM_NAMES, ps_c, ps_g):
            np.testing.assert_allclose(b.cpu().numpy(), a.numpy(), rtol=2e-6, atol=1e-7, err_msg=name)
    assert list(opt.steps[:6]) == [5] * 6 and list(opt.steps[6:12]) == [3] * 6 and list(opt.steps[12:]) == [5] * 6


def test_training_continues_after_refinement():
    """render -> loss -> backward -> after_train (statistics) -> FusedAdam, a refinement at step 700 that changes every
    sub-model's row count, then the same loop again on the new tensors."""
    fr = syn.make_frame(n_background=20000, n_actors=3, n_per_actor=1500, width=320, height=240, seed=3,
                        actor_shift=np.array([1.0, 0.0, -1.0]))
    dev = torch.device("cuda", 0)
    bg = fr.segments[0].params.to(dev)
    actors = {s.name.replace("object_", ""): s.params.to(dev) for s in fr.segments[1:]}
    poses = [ActorPose(s.name.replace("object_", ""), s.rot, s.center, 21, list(range(85))) for s in fr.segments[1:]]
    # thresholds scaled to this synthetic scene so that every category is populated after a few steps
    rs = refine.RefineSettings(densify_grad_thresh=2e-5, densify_size_thresh=0.05, cull_alpha_thresh=0.05)
    cfg = SceneGraphConfig(use_sky_sphere=False, ssim_lambda=0.0, refine=rs, object_refine=rs, num_train_data=5, refine_record=True)
    model = SceneGraphRasterModel(bg, actors, cfg, poses_at=lambda t: poses).to(dev)
    model.train()
    gt = (torch.rand(fr.camera.height, fr.camera.width, 3, generator=torch.Generator().manual_seed(2)) * 0.5 + 0.25).to(dev)
    opt = FusedAdam(model.optimizer_params())

    def train_step(step):
        model.step = step
        for p in model.parameters():
            p.grad = None
        out = model.get_outputs(fr.camera)
        loss = sum(model.get_loss_dict(out, {"image": gt}).values())
        loss.backward()
        model.after_train(step)
        opt.step(model._holder.grad_arena, present=model.present_submodels())
        return float(loss.detach())

    before = [train_step(695 + i) for i in range(5)]
    counts0 = [sub.num_points for sub in model.all_models.values()]
    model.step = 700
    torch.manual_seed(4)
    model.refinement_after(opt, 700, sync_stats=False)
    counts1 = [sub.num_points for sub in model.all_models.values()]
    recs = [sub.refine_record_dict for sub in model.all_models.values()]
    assert counts1 != counts0 and all(c > 0 for c in counts1), (counts0, counts1)
    assert recs[0]["refine_splits_count"] > 0 and recs[0]["refine_dups_count"] > 0, recs[0]
    for sub, c0, c1, r in zip(model.all_models.values(), counts0, counts1, recs):
        assert sub.xys_grad_norm is None
        for k in PARAM_NAMES:
            p = sub.gauss_params[k]
            assert p.shape[0] == c1 and torch.isfinite(p).all()
    assert opt.arena_elems == sum((p.numel() + 3) // 4 * 4 for ps in model.optimizer_params() for p in ps)
    snap = [sub.gauss_params["means"].detach().clone() for sub in model.all_models.values()]
    after = [train_step(701 + i) for i in range(5)]
    torch.cuda.synchronize()
    assert all(np.isfinite(x) for x in before + after), (before, after)
    for sub, m0 in zip(model.all_models.values(), snap):  # the optimizer steps the NEW tensors
        m1 = sub.gauss_params["means"].detach()
        assert torch.isfinite(m1).all() and not torch.equal(m1, m0)
    bgm = model.all_models["background"]
    assert bgm.vis_counts is not None and bgm.vis_counts.shape[0] == counts1[0]  # statistics restarted at the new size


def test_cuda_matches_golden_refine():
    """The committed known-answer vectors (tests/golden/refine_case.npz, generated by tests/golden/make_golden_refine.py)."""
    from tests.test_refine import check_against_golden, product_on_golden
    res = product_on_golden(torch.device("cuda", 0))
    torch.cuda.synchronize()
    check_against_golden(*res, tol=1e-5)


def _alternating_scene(full_arena: bool):
    """Background + 3 actors; actor 1 has a box only at even timestamps (an actor is in view for part of the frames)."""
    from street_gaussians_ns_b200.training import TrainStep
    fr = syn.make_frame(n_background=20000, n_actors=3, n_per_actor=1500, width=320, height=240, seed=3,
                        actor_shift=np.array([1.0, 0.0, -1.0]))
    dev = torch.device("cuda", 0)
    bg = fr.segments[0].params.to(dev)
    actors = {s.name.replace("object_", ""): s.params.to(dev) for s in fr.segments[1:]}
    poses = [ActorPose(s.name.replace("object_", ""), s.rot, s.center, 21, list(range(85))) for s in fr.segments[1:]]

    def poses_at(t):
        return poses if int(t) % 2 == 0 else [poses[0], poses[2]]

    cfg = SceneGraphConfig(use_sky_sphere=False, ssim_lambda=0.0, full_gradient_arena=full_arena)
    model = SceneGraphRasterModel(bg, actors, cfg, poses_at=poses_at).to(dev)
    model.train()
    opt = FusedAdam(model.optimizer_params())
    cams = [syn.make_camera(320, 240, time=float(t)) for t in range(2)]
    gt = (torch.rand(240, 320, 3, generator=torch.Generator().manual_seed(2)) * 0.5 + 0.25).to(dev)
    return model, opt, TrainStep(model, opt, refine_every=0), cams, gt


@pytest.mark.parametrize("full_arena", [False, True])
def test_train_step_skips_actors_that_are_not_in_view(full_arena):
    """TrainStep (training.py) over frames in which actor 1 comes and goes: in the frames without it its parameters, moments
    and step count stay untouched (torch.optim.Adam skips parameters without a gradient), in both arena layouts."""
    model, opt, step_fn, cams, gt = _alternating_scene(full_arena)
    a1 = model.all_models["object_1"]
    for it in range(6):
        before = [a1.gauss_params[k].detach().clone() for k in PARAM_NAMES]
        m_before = opt.moment_views(6 * 2 + 0)[0].clone()   # all_models order: background, object_0, object_1, object_2
        losses = step_fn(1000 + it, cams[it % 2], {"image": gt})
        assert all(torch.isfinite(v) for v in losses.values())
        same = all(torch.equal(a1.gauss_params[k].detach(), b) for k, b in zip(PARAM_NAMES, before))
        if it % 2 == 1:
            assert model.present_submodels() == [0, 1, 3]
            assert same and torch.equal(opt.moment_views(12)[0], m_before)
            assert a1.gauss_params["means"].grad is None
        else:
            assert model.present_submodels() == [0, 1, 2, 3]
            assert not same
    assert list(opt.steps[12:18]) == [3] * 6 and list(opt.steps[:12]) == [6] * 12 and list(opt.steps[18:]) == [6] * 6


def test_full_arena_layout_trains_like_the_frame_layout():
    """Same scene, same steps, once with the frame-layout gradient arena and once with the data-parallel (all sub-models)
    layout: the same gradients land at different offsets, the parameters follow the same trajectory (up to the
    summation-order noise of the backward's atomics)."""
    results = []
    for full in (False, True):
        model, opt, step_fn, cams, gt = _alternating_scene(full)
        for it in range(4):
            step_fn(1000 + it, cams[it % 2], {"image": gt})
        torch.cuda.synchronize()
        results.append({n: {k: sub.gauss_params[k].detach().clone() for k in PARAM_NAMES} for n, sub in model.all_models.items()})
        if full:
            sink = model._grad_sink
            assert sink.arena.numel() == opt.arena_elems  # the optimizer's layout
    for name in results[0]:
        for k in PARAM_NAMES:
            a, b = results[0][name][k], results[1][name][k]
            bad = ((a - b).abs() > 1e-4 + 1e-3 * b.abs()).float().mean().item()
            assert bad < 1e-2, (name, k, bad)  # a wrong offset would mismatch (nearly) everything

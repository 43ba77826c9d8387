"""Host logic of training.TrainStep that needs no GPU: the refinement callbacks still fire on a step whose frame has
nothing in view (nerfstudio fires them on the step count, sgn_splatfacto.py:773-781), and the split samples come from a
generator that depends on (seed, step) only -- whatever else consumed the global generator on a replica."""
import torch

from tests.test_refine import build_model


def make_step(**kw):
    from street_gaussians_ns_b200.optim import FusedAdam
    from street_gaussians_ns_b200.training import TrainStep
    model, _ = build_model()
    opt = FusedAdam(model.optimizer_params(), chunk_elems=4096)
    return model, opt, TrainStep(model, opt, **kw)


def test_refinement_generator_depends_on_seed_and_step_only():
    _, _, a = make_step(refine_every=100, refine_seed=3)
    _, _, b = make_step(refine_every=100, refine_seed=3)
    torch.manual_seed(1)
    torch.randn(17)                      # replica a consumed the global generator differently from replica b
    xa = torch.randn((5, 3), generator=a._refine_generator(600))
    torch.manual_seed(2)
    xb = torch.randn((5, 3), generator=b._refine_generator(600))
    assert torch.equal(xa, xb)
    assert not torch.equal(xa, torch.randn((5, 3), generator=a._refine_generator(700)))
    _, _, c = make_step(refine_every=100, refine_seed=4)
    assert not torch.equal(xa, torch.randn((5, 3), generator=c._refine_generator(600)))


def test_refinement_runs_on_a_step_without_anything_in_view(monkeypatch):
    model, opt, step_fn = make_step(refine_every=100)
    calls = []
    monkeypatch.setattr(model, "get_outputs", lambda camera: {"rgb": torch.zeros(4, 4, 3)})
    monkeypatch.setattr(model, "get_loss_dict", lambda out, batch: {"main_loss": torch.zeros(())})  # no graph: early-out frame
    monkeypatch.setattr(model, "refinement_after", lambda o, step, **kw: calls.append((step, kw.get("due_only"))))
    before = opt.step_count
    step_fn(600, camera=None, batch={})
    step_fn(601, camera=None, batch={})
    assert calls == [(600, False)]        # fired on the step count; 601 is not due
    assert opt.step_count == before       # no optimizer step without a gradient


def test_per_submodel_cadence_when_no_global_one_is_given(monkeypatch):
    model, opt, step_fn = make_step(refine_every=None)
    model.config.refine.refine_every, model.config.object_refine.refine_every = 100, 150
    calls = []
    monkeypatch.setattr(model, "get_outputs", lambda camera: {})
    monkeypatch.setattr(model, "get_loss_dict", lambda out, batch: {"main_loss": torch.zeros(())})
    monkeypatch.setattr(model, "refinement_after", lambda o, step, **kw: calls.append((step, kw.get("due_only"))))
    for s in (100, 125, 150, 300):
        step_fn(s, camera=None, batch={})
    assert calls == [(100, True), (150, True), (300, True)]


def test_warm_up_refinement_exercises_every_branch_and_leaves_the_model_alone(monkeypatch):
    from tests.host_harness import load_refine_harness
    from street_gaussians_ns_b200 import refine, training
    harness = load_refine_harness()
    monkeypatch.setattr(refine, "_backend", lambda: harness)
    monkeypatch.setattr(refine, "_require_cuda", lambda t, what: None)
    model, _ = build_model()
    model.config.num_train_data = 425
    before = [p.detach().clone() for p in model.parameters()]
    info = training.warm_up_refinement(model, rows=(4000, 1000, 1000))
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert refine.phase(model.config.refine, info["step"], 425)[0]               # a densifying step was picked
    assert all(a != b for a, b in zip(info["rows_before"], info["rows_after"]))  # every sub-model was re-laid out
    for rec in info["records"]:  # splits, duplicates and culls all happened: the kernels' branches and the sampler ran
        assert rec["refine_splits_count"] > 0 and rec["refine_dups_count"] > 0 and rec["refine_culls_alpha_count"] > 0

"""Refinement (split / duplicate / cull / opacity reset) on the CPU.

There is no GPU in the build container, so the CUDA kernels themselves are exercised by tests/test_gpu_refine.py on the
B200.  Here the SAME row rules (csrc/sgn_refine_rules.cuh, the bodies of those kernels) are compiled with g++ into a
host harness (tests/host_harness/refine_host.cpp) that stands in for the two C-ABI entry points, and the product's
host side (refine.py, model.refinement_after, FusedAdam.rebuild) is driven through it and compared with the torch
restatement of the reference (oracle/oracle_refine.py; street_gaussians_ns/sgn_splatfacto.py:550-720):

  * a hand-built six-row case with known answers pins the oracle AND the rules (every category once, including the
    reference's quirk that a split row whose shrunk scale drops under the size threshold is also duplicated);
  * random sub-models through every phase of the schedule: tensors, row order, Adam moments, record counters;
  * the model-level two-phase path with FusedAdam arenas and with the reference's per-group torch.optim.Adam form.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_refine as orc
from street_gaussians_ns_b200 import _lib, refine
from street_gaussians_ns_b200.scene import PARAM_NAMES, GaussianSet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    from tests.host_harness import load_refine_harness
    return load_refine_harness()


@pytest.fixture()
def host_backend(harness, monkeypatch):
    """Route refine.py's two library calls to the host harness and accept CPU tensors (tests only)."""
    monkeypatch.setattr(refine, "_backend", lambda: harness)
    monkeypatch.setattr(refine, "_require_cuda", lambda t, what: None)
    return harness


def test_product_refuses_cpu_tensors():
    cfg = refine.make_config(refine.RefineSettings(), 1000, (64, 48), True)
    z = torch.zeros(4, 3)
    with pytest.raises(_lib.SgnError, match="no CPU path"):
        refine.plan_submodel(z, torch.zeros(4, 1), torch.zeros(4), torch.ones(4), torch.zeros(4), cfg)


def test_phase_schedule():
    s = refine.RefineSettings()  # refine_every 100, reset_alpha_every 30 -> reset interval 3000
    assert refine.phase(s, 500, 100) == (False, False, False)          # warm-up (:552-553)
    assert refine.phase(s, 600, 100) == (True, False, False)           # 600 % 3000 > 100 + 100
    assert refine.phase(s, 3100, 100) == (False, False, True)          # right after a reset boundary: reset, no densify
    assert refine.phase(s, 3200, 100) == (False, False, False)         # not every image seen yet since the reset
    assert refine.phase(s, 3300, 100) == (True, False, False)
    assert refine.phase(s, 25000, 100) == (False, True, False)         # past stop_split_at: cull only (:620-621)
    s2 = refine.RefineSettings(continue_cull_post_densification=False)
    assert refine.phase(s2, 25000, 100) == (False, False, False)
    assert abs(refine.opacity_reset_logit(s) - float(np.log(0.04 / 0.96))) < 1e-6


def make_state(n, F, seed, with_moments=True):
    g = torch.Generator().manual_seed(seed)
    p = {
        "means": torch.randn(n, 3, generator=g) * 5,
        # log-scales around the size thresholds (0.01 densify, 0.2 cull)
        "scales": torch.log(torch.exp(torch.randn(n, 3, generator=g) * 1.5 - 4.0)),
        "quats": torch.randn(n, 4, generator=g),
        "features_dc": torch.randn(n, F, 3, generator=g),
        "features_rest": torch.randn(n, 15, 3, generator=g),
        "opacities": torch.randn(n, 1, generator=g) * 2.5 - 1.0,
    }
    moments = None
    if with_moments:
        moments = {k: (torch.randn(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in p.items()}
    vis = torch.randint(1, 9, (n,), generator=g).float()
    xgn = torch.rand(n, generator=g) * vis * 2.5e-6   # avg * 0.5 * 320 straddles densify_grad_thresh 2e-4
    m2d = torch.rand(n, generator=g) * 0.2            # straddles split 0.05 / cull 0.15
    return orc.SubModelState(p, moments, xgn, vis, m2d)


def clone_state(st):
    return orc.SubModelState({k: v.clone() for k, v in st.params.items()},
                             None if st.moments is None else {k: (a.clone(), b.clone()) for k, (a, b) in st.moments.items()},
                             st.xys_grad_norm.clone(), st.vis_counts.clone(), st.max_2Dsize.clone())


def to_settings(cfg: orc.RefineConfig) -> refine.RefineSettings:
    return refine.RefineSettings(**{k: getattr(cfg, k) for k in refine.RefineSettings.__dataclass_fields__})


def run_product(st, cfg, step, size, ntrain, seed):
    params = [st.params[k].contiguous() for k in PARAM_NAMES]
    moments = None if st.moments is None else [tuple(x.contiguous() for x in st.moments[k]) for k in PARAM_NAMES]
    g = torch.Generator().manual_seed(seed)
    return refine.refine_tensors(params, moments, (st.xys_grad_norm, st.vis_counts, st.max_2Dsize), to_settings(cfg), step, size,
                                 ntrain, generator=g)


def run_oracle(st, cfg, step, size, ntrain, seed):
    g = torch.Generator().manual_seed(seed)
    rec = orc.refinement_after(st, cfg, step, size, ntrain, randn=lambda k: torch.randn((k, 3), generator=g))
    return st, rec


def assert_same(new_p, new_m, ost, label):
    for k, name in enumerate(PARAM_NAMES):
        a, b = new_p[k], ost.params[name]
        assert a.shape == b.shape, (label, name, a.shape, b.shape)
        if name in ("means", "scales"):  # exp/log (and the rotation) go through libm here, torch's vector math there
            torch.testing.assert_close(a, b, rtol=2e-6, atol=2e-6, msg=lambda m: f"{label} {name}: {m}")
        else:
            assert torch.equal(a, b), (label, name)
        if ost.moments is not None:
            assert torch.equal(new_m[k][0], ost.moments[name][0]), (label, name, "exp_avg")
            assert torch.equal(new_m[k][1], ost.moments[name][1]), (label, name, "exp_avg_sq")


def test_hand_built_case_pins_oracle_and_rules(host_backend):
    """Six rows, one per category; size (100, 80) -> avg = norm / vis * 0.5 * 100."""
    cfg = orc.RefineConfig(cull_alpha_thresh=0.1, cull_scale_thresh=0.5)
    big, small, mid = np.log(0.05), np.log(0.004), np.log(0.012)
    scales = torch.tensor([[big, small, small],      # A: high grad, large        -> split (2 samples), row removed
                           [small, small, small],    # B: high grad, small        -> duplicated
                           [small, small, small],    # C: low grad, transparent   -> culled (alpha)
                           [np.log(0.9), small, small],  # D: low grad, huge      -> culled (too big) once cull_big
                           [mid, small, small],      # E: high grad, 0.012 > 0.01 -> split; 0.012/1.6 = 0.0075 <= 0.01 -> ALSO duplicated
                           [small, small, small]],   # F: low grad                -> kept
                          dtype=torch.float32)
    n = 6
    p = {"means": torch.arange(18, dtype=torch.float32).view(6, 3), "scales": scales,
         "quats": torch.tensor([[1.0, 0, 0, 0]] * n), "features_dc": torch.arange(6.0).view(6, 1, 1).repeat(1, 1, 3),
         "features_rest": torch.zeros(n, 15, 3), "opacities": torch.tensor([[2.0], [2.0], [-5.0], [2.0], [2.0], [2.0]])}
    vis = torch.full((n,), 2.0)
    xgn = torch.tensor([1.0, 1.0, 0.0, 0.0, 1.0, 0.0]) * 1e-4   # avg = 1e-4 / 2 * 50 = 2.5e-3 > 2e-4 for A, B, E
    m2d = torch.zeros(n)
    step, size, ntrain = 3400, (100, 80), 10                       # densify, cull_big (3400 > 3000), screen size on (< 4000)
    st = orc.SubModelState({k: v.clone() for k, v in p.items()}, None, xgn, vis, m2d)
    new_p, _, plan = run_product(clone_state(st), cfg, step, size, ntrain, seed=5)
    ost, rec = run_oracle(st, cfg, step, size, ntrain, seed=5)
    # survivors B, F; then samples [A0, E0, A1, E1]; then duplicates [B, E']
    assert plan.totals == [2, 2, 2, 2] and plan.out_rows == 8
    ids = ost.params["features_dc"][:, 0, 0].tolist()
    assert ids == [1.0, 5.0, 0.0, 4.0, 0.0, 4.0, 1.0, 4.0]
    assert rec["refine_splits_count"] == 2 and rec["refine_dups_count"] == 2 and rec["high_grads_count"] == 3
    assert rec["refine_culls_alpha_count"] == 1 and rec["refine_culls_toobigs_count"] == 1
    assert_same(new_p, None, ost, "hand")
    # identity rotation: a sample sits at mean + exp(scale) * z, new scales are log(exp(s)/1.6), the duplicate of E is shrunk
    z = torch.randn((4, 3), generator=torch.Generator().manual_seed(5))
    torch.testing.assert_close(new_p[0][2], p["means"][0] + torch.exp(scales[0]) * z[0], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(new_p[0][5], p["means"][4] + torch.exp(scales[4]) * z[3], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(torch.exp(new_p[1][3]), torch.exp(scales[4]) / 1.6, rtol=1e-5, atol=0)
    torch.testing.assert_close(torch.exp(new_p[1][7]), torch.exp(scales[4]) / 1.6, rtol=1e-5, atol=0)
    assert torch.equal(new_p[1][6], scales[1])
    got = plan.record()
    assert got["refine_splits_count"] == 2 and got["refine_dups_count"] == 2 and got["high_grads_count"] == 3
    assert got["refine_culls_alpha_count"] == 1 and got["refine_culls_toobigs_count"] == 1


PHASES = [
    # (label, step, config overrides): with reset interval 3000 and num_train_data 50
    ("densify+screen", 700, {}),                        # < stop_screen_size_at, not yet cull_big
    ("densify+screen+big", 3400, {}),                   # cull_big and screen-size culling
    ("densify", 4400, {}),                              # past stop_screen_size_at
    ("densify+3samples", 7400, {"n_split_samples": 3}),
    ("reset-only", 3100, {}),                           # opacity reset, no structural change
    ("cull-only", 25000, {}),                           # past stop_split_at
    ("nothing", 3130, {}),                              # 130 <= num_train_data + refine_every: not every image seen since the reset
]


@pytest.mark.parametrize("label,step,over", PHASES, ids=[p[0] for p in PHASES])
@pytest.mark.parametrize("F,moments", [(1, True), (5, True), (1, False)])
def test_random_submodel_matches_oracle(host_backend, label, step, over, F, moments):
    cfg = orc.RefineConfig(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2, **over)
    st = make_state(2500 + 7 * F, F, seed=11 + F, with_moments=moments)
    size, ntrain = (240, 320), 50
    new_p, new_m, plan = run_product(clone_state(st), cfg, step, size, ntrain, seed=3)
    ost, rec = run_oracle(st, cfg, step, size, ntrain, seed=3)
    assert_same(new_p, new_m, ost, label)
    if label.startswith("densify"):
        assert plan.totals[3] > 50 and plan.totals[2] > 50 and plan.totals[0] < plan.n  # every category is populated
        got = plan.record()
        for k in ("high_grads_count", "refine_splits_count", "refine_dups_count", "refine_culls_alpha_count"):
            assert got[k] == rec[k], (k, got, rec)
    if label in ("nothing", "reset-only"):
        assert plan is None  # no decision pass at all: the tensors were not rebuilt
    if label == "cull-only":
        assert plan.totals[1] == plan.totals[2] == plan.totals[3] == 0 and 0 < plan.totals[0] < plan.n


def test_no_statistics_means_no_refinement(host_backend):
    st = make_state(100, 1, 0)
    params = [st.params[k] for k in PARAM_NAMES]
    new_p, _, plan = refine.refine_tensors(params, None, (None, None, None), refine.RefineSettings(), 700, (64, 48), 10)
    assert plan is None and all(a is b for a, b in zip(new_p, params))


def test_everything_culled_and_empty_submodel(host_backend):
    cfg = orc.RefineConfig(stop_split_at=25000)
    st = make_state(64, 1, 3)
    st.params["opacities"].fill_(-20.0)  # sigmoid < any threshold: every row, sample and duplicate is culled
    new_p, new_m, plan = run_product(clone_state(st), cfg, 700, (64, 48), 10, seed=1)
    assert plan.out_rows == 0 and all(t.shape[0] == 0 for t in new_p) and all(m.shape[0] == 0 for m, _ in new_m)
    ost, _ = run_oracle(st, cfg, 700, (64, 48), 10, seed=1)
    assert ost.params["means"].shape[0] == 0
    # an empty sub-model (0 rows) passes through
    empty = [t[:0].contiguous() for t in new_p]
    z = torch.zeros(0)
    out, _, plan0 = refine.refine_tensors(empty, None, (z, z, z), to_settings(cfg), 700, (64, 48), 10)
    assert plan0.out_rows == 0 and not plan0.changed and all(a is b for a, b in zip(out, empty))


# --------------------------------------------------------------------------------------------------
# model level: two-phase path, FusedAdam arenas, the reference's per-group torch.optim.Adam form
# --------------------------------------------------------------------------------------------------
def build_model(seed=0):
    from street_gaussians_ns_b200.model import SceneGraphConfig, SceneGraphRasterModel
    states = [make_state(900, 1, seed), make_state(300, 5, seed + 1), make_state(200, 5, seed + 2)]
    sets = [GaussianSet(*[s.params[k].clone() for k in PARAM_NAMES]) for s in states]
    cfg = SceneGraphConfig(use_sky_sphere=False, num_train_data=50)
    model = SceneGraphRasterModel(sets[0], {"a": sets[1], "b": sets[2]}, cfg)
    model.train()
    for sub, s in zip(model.all_models.values(), states):
        d = sub.__dict__
        d["xys_grad_norm"], d["vis_counts"], d["max_2Dsize"] = s.xys_grad_norm.clone(), s.vis_counts.clone(), s.max_2Dsize.clone()
        d["last_size"] = (240, 320)
    return model, states


def oracle_cfg_for(name):
    return orc.RefineConfig(stop_split_at=25000, cull_scale_thresh=0.2, cull_alpha_thresh=0.02 if name == "background" else 0.005)


@pytest.mark.parametrize("step", [700, 3100, 3400])
def test_model_refinement_with_fused_adam(host_backend, step):
    from street_gaussians_ns_b200.optim import FusedAdam
    model, states = build_model()
    model.step = step
    opt = FusedAdam(model.optimizer_params(), chunk_elems=4096)
    g = torch.Generator().manual_seed(9)
    opt.exp_avg.copy_(torch.randn(opt.arena_elems, generator=g))
    opt.exp_avg_sq.copy_(torch.rand(opt.arena_elems, generator=g))
    opt.steps[:] = np.arange(len(opt.steps)) + 1
    # the oracle's view of the same optimizer state
    for i, s in enumerate(states):
        s.moments = {k: tuple(x.clone() for x in opt.moment_views(6 * i + j)) for j, k in enumerate(PARAM_NAMES)}
    gen = torch.Generator().manual_seed(21)
    model.refinement_after(opt, step, generator=gen, sync_stats=False)
    ogen = torch.Generator().manual_seed(21)
    for i, (name, s) in enumerate(zip(model.all_models.keys(), states)):
        orc.refinement_after(s, oracle_cfg_for(name), step, (240, 320), 50, randn=lambda k: torch.randn((k, 3), generator=ogen))
        sub = model.all_models[name]
        new_p = [sub.gauss_params[k].data for k in PARAM_NAMES]
        assert all(isinstance(sub.gauss_params[k], torch.nn.Parameter) and sub.gauss_params[k].requires_grad for k in PARAM_NAMES)
        assert_same(new_p, [opt.moment_views(6 * i + j) for j in range(6)], s, f"{name}@{step}")
        assert sub.xys_grad_norm is None and sub.vis_counts is None and sub.max_2Dsize is None
        # the optimizer now points at the model's tensors, and kept its step counts
        for j, k in enumerate(PARAM_NAMES):
            assert int(opt.table["param"][6 * i + j]) == sub.gauss_params[k].data_ptr()
            assert int(opt.table["numel"][6 * i + j]) == sub.gauss_params[k].numel()
    assert list(opt.steps) == list(np.arange(len(opt.steps)) + 1)
    assert opt.exp_avg.numel() == opt.arena_elems == sum((p.numel() + 3) // 4 * 4 for ps in model.optimizer_params() for p in ps)


def test_model_refinement_with_reference_style_adam_groups(host_backend):
    step = 3400
    model, states = build_model(seed=4)
    model.step = step
    params = model.optimizer_params()
    groups = {k: torch.optim.Adam([ps[j] for ps in params], lr=1e-3, eps=1e-15) for j, k in enumerate(PARAM_NAMES)}
    g = torch.Generator().manual_seed(2)
    for ps in params:  # one optimizer step so that the state exists
        for p in ps:
            p.grad = torch.randn(p.shape, generator=g)
    for o in groups.values():
        o.step()
        o.zero_grad(set_to_none=True)
    for i, s in enumerate(states):
        s.params = {k: params[i][j].data.clone() for j, k in enumerate(PARAM_NAMES)}
        s.moments = {k: (groups[k].state[params[i][j]]["exp_avg"].clone(), groups[k].state[params[i][j]]["exp_avg_sq"].clone())
                     for j, k in enumerate(PARAM_NAMES)}

    class Optimizers:  # nerfstudio's container: .optimizers[name]
        optimizers = groups

    model.refinement_after(Optimizers(), step, generator=torch.Generator().manual_seed(8), sync_stats=False)
    ogen = torch.Generator().manual_seed(8)
    for i, (name, s) in enumerate(zip(model.all_models.keys(), states)):
        orc.refinement_after(s, oracle_cfg_for(name), step, (240, 320), 50, randn=lambda k: torch.randn((k, 3), generator=ogen))
        sub = model.all_models[name]
        moments = []
        for j, k in enumerate(PARAM_NAMES):
            p = groups[k].param_groups[0]["params"][i]
            assert p is sub.gauss_params[k]                      # the group now holds the new parameter
            st = groups[k].state[p]
            assert float(st["step"]) == 1.0                      # param_state moved as a whole (:459-476)
            moments.append((st["exp_avg"], st["exp_avg_sq"]))
        assert_same([sub.gauss_params[k].data for k in PARAM_NAMES], moments, s, name)
        assert len(groups["means"].state) == len(states)         # no stale entries


def test_fused_adam_step_table_skips_absent_submodels():
    """torch.optim.Adam skips parameters without a gradient (no decay, no step): the table of a step only lists the
    sub-models present in the frame's arena, with gradient offsets in THAT arena, and per-tensor bias corrections."""
    from street_gaussians_ns_b200.optim import FusedAdam
    params = [[torch.zeros(n, *shape) for shape in ((3,), (3,), (4,), (F, 3), (15, 3), (1,))] for n, F in ((10, 1), (7, 5), (5, 5))]
    opt = FusedAdam(params, chunk_elems=64)
    t_all = opt.step_table()
    assert len(t_all) == 18 and np.array_equal(t_all["grad_offset"], t_all["arena_offset"])
    t = opt.step_table(present=[0, 2]).copy()
    assert len(t) == 12
    assert list(opt.steps) == [2] * 6 + [1] * 6 + [2] * 6
    sizes = [(p.numel() + 3) // 4 * 4 for p in params[0] + params[2]]
    assert list(t["grad_offset"]) == list(np.concatenate([[0], np.cumsum(sizes)[:-1]]))
    assert list(t["arena_offset"][6:]) == list(opt.offsets[12:])            # moments stay where the optimizer keeps them
    chunks = [(p.numel() + 63) // 64 for p in params[0] + params[2]]
    assert list(t["chunk0"]) == list(np.concatenate([[0], np.cumsum(chunks)[:-1]]))
    np.testing.assert_allclose(t["step_size"][0], 1.6e-4 / (1 - 0.9 ** 2), rtol=1e-6)
    t1 = opt.step_table(present=[1])
    np.testing.assert_allclose(t1["step_size"][0], 1.6e-4 / (1 - 0.9 ** 2), rtol=1e-6)   # its second step only now
    np.testing.assert_allclose(t1["sqrt_bc2"][0], np.sqrt(1 - 0.999 ** 2), rtol=1e-6)


# --------------------------------------------------------------------------------------------------
# committed known-answer vectors (tests/golden/refine_case.npz)
# --------------------------------------------------------------------------------------------------
def product_on_golden(device):
    """The product path (refine.plan_submodel / apply_plan) on the fixture's inputs, with the fixture's split samples."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_refine as mg
    gold = np.load(mg.OUT)
    t = lambda k: torch.from_numpy(gold[k].copy()).to(device)  # noqa: E731
    params = [t("in_" + k) for k in PARAM_NAMES]
    moments = [(t("in_m_" + k), t("in_v_" + k)) for k in PARAM_NAMES]
    settings = refine.RefineSettings(**mg.CONFIG)
    cfg = refine.make_config(settings, mg.STEP, mg.SIZE, True)
    plan = refine.plan_submodel(params[1], params[5], t("xys_grad_norm"), t("vis_counts"), t("max_2Dsize"), cfg)
    plan.samples = t("samples")[: cfg.n_split_samples * plan.totals[3]].contiguous()
    new = [torch.empty((plan.out_rows,) + tuple(p.shape[1:]), device=device) for p in params]
    new_m = [(torch.empty_like(a), torch.empty_like(a)) for a in new]
    refine.apply_plan(plan, params, new, moments, new_m)
    return gold, plan, new, new_m


def check_against_golden(gold, plan, new, new_m, tol):
    for k, name in enumerate(PARAM_NAMES):
        want = gold["out_" + name]
        got = new[k].cpu().numpy()
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if name in ("means", "scales"):
            np.testing.assert_allclose(got, want, rtol=tol, atol=tol, err_msg=name)
        else:
            np.testing.assert_array_equal(got, want, err_msg=name)
        np.testing.assert_array_equal(new_m[k][0].cpu().numpy(), gold["out_m_" + name], err_msg=name + " exp_avg")
        np.testing.assert_array_equal(new_m[k][1].cpu().numpy(), gold["out_v_" + name], err_msg=name + " exp_avg_sq")
    rec = plan.record()
    got = [rec["high_grads_count"], rec["refine_splits_count"], rec["refine_dups_count"], rec["refine_culls_alpha_count"]]
    assert got == [int(x) for x in gold["counts"][:4]]


def test_oracle_reproduces_golden_refine():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_refine as mg
    gold = np.load(mg.OUT)
    out = mg.run_oracle({k: gold[k] for k in gold.files})
    for k, v in out.items():
        if k in ("out_means", "out_scales"):
            np.testing.assert_allclose(v, gold[k], rtol=1e-6, atol=1e-6, err_msg=k)  # vector exp/log may differ by an ulp across hosts
        else:
            np.testing.assert_array_equal(v, gold[k], err_msg=k)


def test_rules_match_golden_refine(host_backend):
    check_against_golden(*product_on_golden(torch.device("cpu")), tol=2e-6)


def test_checkpoint_with_a_different_number_of_gaussians_loads():
    """Refinement changes the row counts; a checkpoint taken later must load into a freshly built model
    (sgn_splatfacto.py:425-438, scene graph :393-401), including the pre-ParameterDict key names."""
    big, _ = build_model(seed=0)
    small, _ = build_model(seed=5)
    for sub in small.all_models.values():   # a model "before refinement": fewer rows everywhere
        for k in PARAM_NAMES:
            sub.gauss_params[k] = torch.nn.Parameter(sub.gauss_params[k].data[:50].clone())
    ckpt = {k: v.clone() for k, v in big.state_dict().items()}
    assert "all_models.background.gauss_params.means" in ckpt and "all_models.object_a.gauss_params.features_dc" in ckpt
    small.load_state_dict(ckpt)
    for name in big.all_models:
        for k in PARAM_NAMES:
            assert torch.equal(small.all_models[name].gauss_params[k], big.all_models[name].gauss_params[k]), (name, k)
            assert isinstance(small.all_models[name].gauss_params[k], torch.nn.Parameter)
        assert small.all_models[name].xys_grad_norm is None
    # old key names (means, scales, ... directly on the sub-model)
    sub = small.all_models["object_b"]
    legacy = {k: torch.full_like(big.all_models["object_a"].gauss_params[k], 0.5) for k in PARAM_NAMES}
    sub.load_state_dict(legacy)
    assert sub.num_points == big.all_models["object_a"].num_points and float(sub.gauss_params["quats"].detach().min()) == 0.5


def test_fused_adam_relayout_keeps_untouched_submodels(host_backend):
    """One actor has no statistics (never in view since the last refinement): its tensors stay, but its moments move to
    new offsets inside the rebuilt arenas because the sub-models before it changed size."""
    from street_gaussians_ns_b200.optim import FusedAdam
    step = 3400
    model, states = build_model(seed=2)
    model.step = step
    quiet = model.all_models["object_a"]
    d = quiet.__dict__
    d["xys_grad_norm"] = d["vis_counts"] = d["max_2Dsize"] = None
    opt = FusedAdam(model.optimizer_params(), chunk_elems=4096)
    g = torch.Generator().manual_seed(3)
    opt.exp_avg.copy_(torch.randn(opt.arena_elems, generator=g))
    opt.exp_avg_sq.copy_(torch.rand(opt.arena_elems, generator=g))
    before_p = [quiet.gauss_params[k].data.clone() for k in PARAM_NAMES]
    before_m = [tuple(x.clone() for x in opt.moment_views(6 + j)) for j in range(6)]
    old_offset = int(opt.offsets[6])
    ptrs = [quiet.gauss_params[k].data_ptr() for k in PARAM_NAMES]
    model.refinement_after(opt, step, generator=torch.Generator().manual_seed(1), sync_stats=False)
    assert model.all_models["background"].num_points != states[0].params["means"].shape[0]  # the arena before it changed
    assert int(opt.offsets[6]) != old_offset
    for j, k in enumerate(PARAM_NAMES):
        assert quiet.gauss_params[k].data_ptr() == ptrs[j] and torch.equal(quiet.gauss_params[k].data, before_p[j])
        m, v = opt.moment_views(6 + j)
        assert torch.equal(m, before_m[j][0]) and torch.equal(v, before_m[j][1]), k
        assert int(opt.table["param"][6 + j]) == ptrs[j]


@pytest.mark.parametrize("F,rest,samps", [(8, 15, 4), (1, 0, 2), (3, 3, 1)])
def test_row_widths_and_sample_counts(host_backend, F, rest, samps):
    """Widest Fourier basis (SGN_MAX_FOURIER = 8), sh_degree 0 (features_rest has zero columns), sh_degree 1, and 1 / 2 / 4
    split samples: the per-element rebuild walks whatever column layout the six tensors have."""
    st = make_state(700, F, seed=40 + F)
    st.params["features_rest"] = st.params["features_rest"][:, :rest].contiguous()
    st.moments["features_rest"] = tuple(x[:, :rest].contiguous() for x in st.moments["features_rest"])
    cfg = orc.RefineConfig(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2, n_split_samples=samps)
    new_p, new_m, plan = run_product(clone_state(st), cfg, 3400, (240, 320), 50, seed=6)
    ost, _ = run_oracle(st, cfg, 3400, (240, 320), 50, seed=6)
    assert plan.totals[3] > 10 and plan.out_rows == ost.params["means"].shape[0]
    assert_same(new_p, new_m, ost, f"F={F} rest={rest} samps={samps}")


def test_special_values_follow_torch_semantics(host_backend):
    """Rows with never-seen statistics (vis_counts = 0 -> 0/0 and x/0), infinite / NaN statistics, overflowing and
    underflowing scales, saturated opacities: every comparison must come out as torch's does (NaN compares false)."""
    n = 64
    st = make_state(n, 1, seed=8)
    inf, nan = float("inf"), float("nan")
    st.vis_counts[:8] = 0.0                     # never visible: xys_grad_norm / 0
    st.xys_grad_norm[:4] = 0.0                  # 0 / 0 = NaN  -> not a high gradient
    st.xys_grad_norm[8:10] = inf
    st.xys_grad_norm[10:12] = nan
    st.max_2Dsize[12:14] = nan
    st.max_2Dsize[14:16] = inf
    st.params["scales"][16:18] = 100.0          # exp overflows to inf
    st.params["scales"][18:20] = -120.0         # exp underflows to 0 (log(0 / 1.6) = -inf for a split row)
    st.params["opacities"][20:22] = 200.0       # sigmoid == 1
    st.params["opacities"][22:24] = -200.0      # sigmoid == 0
    st.params["opacities"][24] = nan
    cfg = orc.RefineConfig(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2)
    for step in (700, 3400, 25000):
        new_p, new_m, plan = run_product(clone_state(st), cfg, step, (240, 320), 50, seed=2)
        ost, _ = run_oracle(clone_state(st), cfg, step, (240, 320), 50, seed=2)
        for k, name in enumerate(PARAM_NAMES):
            a, b = new_p[k], ost.params[name]
            assert a.shape == b.shape, (step, name, a.shape, b.shape)
            if name in ("means", "scales"):
                torch.testing.assert_close(a, b, rtol=2e-6, atol=2e-6, equal_nan=True)
            else:
                assert torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0)), (step, name)


@pytest.mark.parametrize("n", [0, 1, 2, 3])
def test_tiny_submodels(host_backend, n):
    st = make_state(max(n, 1), 5, seed=50 + n)
    if n == 0:
        st = orc.SubModelState({k: v[:0].contiguous() for k, v in st.params.items()},
                               {k: (a[:0].contiguous(), b[:0].contiguous()) for k, (a, b) in st.moments.items()},
                               st.xys_grad_norm[:0], st.vis_counts[:0], st.max_2Dsize[:0])
    cfg = orc.RefineConfig(stop_split_at=25000, cull_alpha_thresh=0.02, cull_scale_thresh=0.2)
    new_p, new_m, plan = run_product(clone_state(st), cfg, 3400, (240, 320), 50, seed=1)
    ost, _ = run_oracle(st, cfg, 3400, (240, 320), 50, seed=1)
    assert_same(new_p, new_m, ost, f"n={n}")

"""CPU tests that PIN THE ORACLE (no GPU): known answers derived from the reference text, the
independent torch/autograd restatement in float64, and closed forms.  The reference has no tests
or golden vectors of its own for this path -- parity unpinned at the gsplat boundary."""
import math

import numpy as np
import pytest
import torch

import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200.scene import (CLS_BACKGROUND, Camera, Frame, GaussianSet, Segment, idft_basis,
                                            quaternion_from_matrix)
from oracle import oracle_c, oracle_torch


def rel_l2(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a)


# ------------------------------------------------------------------ known answers from the reference text
def test_idft_golden():
    # values computed from sgn_splatfacto_scene_graph.py:420-433
    np.testing.assert_allclose(idft_basis(0.25, 5), [1, 0.587785, 0.809017, 0.951057, 0.309017], atol=2e-6)
    np.testing.assert_allclose(idft_basis(1.0, 5), [1, 0.587785, -0.809017, -0.951057, 0.309017], atol=2e-6)
    np.testing.assert_allclose(idft_basis(0.0, 1), [1.0], atol=0)


def test_viewmat_identity_and_flip():
    cam = syn.make_camera(64, 48)
    vm = cam.viewmat()
    np.testing.assert_array_equal(vm[:, :3], np.diag([1.0, -1.0, -1.0]).astype(np.float32))
    np.testing.assert_array_equal(vm[:, 3], np.zeros(3, np.float32))
    # translated + yawed camera: W2C * C2W == diag(1,-1,-1) flip of identity
    y = 0.3
    R = np.array([[math.cos(y), 0, math.sin(y)], [0, 1, 0], [-math.sin(y), 0, math.cos(y)]])
    c2w = np.concatenate([R, np.array([[1.0], [2.0], [3.0]])], 1)
    vm = syn.make_camera(64, 48, c2w=c2w).viewmat().astype(np.float64)
    p_cam_gl = np.array([0.2, -0.1, -4.0])  # a point in the OpenGL camera frame
    p_world = R @ p_cam_gl + np.array([1.0, 2.0, 3.0])
    p_cv = vm[:, :3] @ p_world + vm[:, 3]
    np.testing.assert_allclose(p_cv, [0.2, 0.1, 4.0], atol=1e-6)


def test_quaternion_from_matrix():
    for yaw in (-0.2, 0.0, 0.13, 2.5):
        c, s = math.cos(yaw), math.sin(yaw)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        q = quaternion_from_matrix(R)
        assert q[0] >= 0
        np.testing.assert_allclose(np.linalg.norm(q), 1.0, atol=1e-12)
        Rq = oracle_torch.quat_to_rotmat(torch.from_numpy(q)).numpy()
        np.testing.assert_allclose(Rq, R, atol=1e-12)


def test_expf_spec_accuracy_and_agreement():
    x = np.linspace(-12, 6, 4001).astype(np.float32)
    c = oracle_c.expf_spec(x)
    t = oracle_torch.expf_spec(torch.from_numpy(x)).numpy()
    np.testing.assert_array_equal(c, t)  # same operation sequence -> same bits
    ref = np.exp(x.astype(np.float64))
    ulp = np.abs(c.astype(np.float64) - ref) / np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert ulp.max() <= 4.0, ulp.max()


def _single_gaussian_frame(sigma=0.05, z=5.0, opac_logit=1.0, W=64, H=48):
    ps = GaussianSet(
        means=torch.tensor([[0.0, 0.0, -z]]), scales=torch.full((1, 3), math.log(sigma)),
        quats=torch.tensor([[1.0, 0.0, 0.0, 0.0]]), features_dc=torch.tensor([[[0.3, -0.2, 0.9]]]),
        features_rest=torch.zeros(1, 15, 3), opacities=torch.tensor([[opac_logit]]))
    return Frame(camera=syn.make_camera(W, H), segments=[Segment(ps, CLS_BACKGROUND)])


def test_single_isotropic_gaussian_closed_form():
    """alpha(px) = o*exp(-r^2 / (2 (sigma^2 f^2/z^2 + 0.3))) on the optical axis."""
    sigma, z = 0.05, 5.0
    fr = _single_gaussian_frame(sigma, z)
    cam = fr.camera
    o = oracle_c.Oracle(fr)
    fw = o.forward(class_renders=False)
    assert fw.radii[0] > 0
    var = (sigma * cam.fx / z) ** 2 + 0.3
    np.testing.assert_allclose(fw.conics[0], [1 / var, 0.0, 1 / var], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(fw.xys[0], [cam.cx, cam.cy], atol=1e-4)
    # gsplat's eigenvalue formula floors the discriminant at 0.1 (Appendix A.3)
    assert fw.radii[0] == math.ceil(3 * math.sqrt(var + math.sqrt(0.1)))
    op = 1 / (1 + math.exp(-1.0))
    H, W = cam.height, cam.width
    jj, ii = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    r2 = (jj - cam.cx) ** 2 + (ii - cam.cy) ** 2
    alpha = np.minimum(0.999, op * np.exp(-0.5 * r2 / var))
    alpha[alpha < 1 / 255.0] = 0
    # only tiles inside the 3-sigma AABB receive the Gaussian
    tb = fw.tile_bbox[0]
    mask = np.zeros((H, W), bool)
    mask[tb[1] * 16: tb[3] * 16, tb[0] * 16: tb[2] * 16] = True
    alpha = np.where(mask, alpha, 0)
    np.testing.assert_allclose(1 - fw.final_T, alpha, atol=2e-6)
    rgb = np.maximum(0.28209479177387814 * np.array([0.3, -0.2, 0.9]) + 0.5, 0)
    np.testing.assert_allclose(fw.img[..., :3], alpha[..., None] * rgb, atol=2e-6)
    np.testing.assert_allclose(fw.img[..., 3], alpha * z, atol=2e-5)


# ------------------------------------------------------------------ C oracle vs torch oracle
@pytest.fixture(scope="module")
def small_scene():
    fr = syn.make_frame(6000, 3, n_per_actor=700, width=160, height=112, seed=3,
                        actor_shift=np.array([2.0, 0.0, -2.0]))
    # bring actors into view: use the inner lanes only
    o = oracle_c.Oracle(fr, alpha_clamp_fwd=0.999, alpha_clamp_bwd=0.999)
    fw = o.forward()
    return fr, o, fw


def test_projection_matches_torch_f32_bitwise_ints(small_scene):
    fr, o, fw = small_scene
    _, cat = oracle_torch.compose(fr, torch.float32, requires_grad=False)
    pr = oracle_torch.project(cat, fr.camera)
    np.testing.assert_array_equal(pr["radii"].numpy(), fw.radii)
    np.testing.assert_array_equal(pr["num_tiles_hit"].numpy(), fw.num_tiles_hit)
    vis = fw.radii > 0
    assert vis.sum() > 1000
    np.testing.assert_allclose(pr["xys"].numpy()[vis], fw.xys[vis], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(pr["depths"].numpy()[vis], fw.depths[vis], rtol=1e-6)
    np.testing.assert_allclose(pr["conics"].numpy()[vis], fw.conics[vis], rtol=2e-4, atol=1e-6)
    rgbs, opac = oracle_torch.colours(cat, fr.camera, 3, 3)
    np.testing.assert_allclose(rgbs.numpy(), fw.rgbs, atol=2e-6)
    np.testing.assert_allclose(opac.numpy(), fw.opac, atol=1e-6)


def test_sort_order_and_bins(small_scene):
    fr, o, fw = small_scene
    assert fw.M == int(fw.num_tiles_hit.sum()) and fw.M > 5000
    tiles_x = (fr.camera.width + 15) // 16
    # every tile list is depth-sorted, ties by index; and contains exactly the AABB members
    for t in range(fw.tile_bins.shape[0]):
        b, e = fw.tile_bins[t]
        ids = fw.sorted_ids[b:e]
        d = fw.depths[ids]
        assert np.all(np.diff(d) >= 0)
        ty, tx = divmod(t, tiles_x)
        bb = fw.tile_bbox
        member = (fw.radii > 0) & (bb[:, 0] <= tx) & (tx < bb[:, 2]) & (bb[:, 1] <= ty) & (ty < bb[:, 3])
        np.testing.assert_array_equal(np.sort(ids), np.nonzero(member)[0])
    assert fw.tile_bins[-1, 1] == fw.M


def test_blend_forward_matches_torch_f64(small_scene):
    fr, o, fw = small_scene
    leaves, out = oracle_torch.render(fr, fw.sorted_ids, fw.tile_bins, alpha_clamp=0.999, dtype=torch.float64)
    ok = fw.fragile == 0
    assert ok.mean() > 0.98
    img = out["img"].detach().numpy()
    assert np.abs(img[ok][:, :3] - fw.img[ok][:, :3]).max() < 2e-5
    # depth channel carries metres (values up to ~90): float32-vs-float64 rounding scales with it
    assert np.abs(img[ok][:, 3] - fw.img[ok][:, 3]).max() < 2e-5 * fw.depths.max()
    assert np.abs(out["alpha"].detach().numpy()[ok] - (1 - fw.final_T)[ok]).max() < 2e-5
    oko = fw.fragile_obj == 0
    assert np.abs(out["object_acc"].detach().numpy()[oko] - (1 - fw.obj_T)[oko]).max() < 2e-5
    okb = fw.fragile_bg == 0
    assert np.abs(out["background_acc"].detach().numpy()[okb] - (1 - fw.bg_T)[okb]).max() < 2e-5
    assert (1 - fw.obj_T).max() > 0.5  # actors are in view


def test_backward_matches_autograd_f64(small_scene):
    """Hand-derived C backward (blend + SH + project + compose) == float64 autograd, consistent clamp."""
    fr, o, fw = small_scene
    H, W = fr.camera.height, fr.camera.width
    g = torch.Generator().manual_seed(11)
    ok = ((fw.fragile == 0) & (fw.fragile_obj == 0) & (fw.fragile_bg == 0)).astype(np.float32)
    v_img = (torch.rand(H, W, 4, generator=g).numpy() * ok[..., None]).astype(np.float32)
    v_img[..., 3] *= 0.05
    v_alpha = (torch.rand(H, W, generator=g).numpy() * ok).astype(np.float32)
    v_obj = (torch.rand(H, W, generator=g).numpy() * ok).astype(np.float32)
    v_bg = (torch.rand(H, W, generator=g).numpy() * ok).astype(np.float32)
    grads, raster = o.backward(fw, v_img, v_alpha, v_obj, v_bg)

    leaves, out = oracle_torch.render(fr, fw.sorted_ids, fw.tile_bins, alpha_clamp=0.999, dtype=torch.float64)
    loss = ((out["img"] * torch.from_numpy(v_img).double()).sum() + (out["alpha"] * torch.from_numpy(v_alpha).double()).sum()
            + (out["object_acc"] * torch.from_numpy(v_obj).double()).sum()
            + (out["background_acc"] * torch.from_numpy(v_bg).double()).sum())
    loss.backward()
    for s, (lf, gr) in enumerate(zip(leaves, grads)):
        for k in ("means", "scales", "quats", "features_dc", "features_rest", "opacities"):
            ref = lf[k].grad.numpy()
            assert np.linalg.norm(ref) > 0, (s, k)
            err = rel_l2(gr[k], ref)
            assert err < 2e-4, (s, k, err)


def test_quirk_clamp_changes_backward_only_when_active():
    """gsplat 0.1.x: forward clamps alpha at 0.999, backward at 0.99 (Appendix A.6).  With opacities
    < 0.99 the two backward modes agree exactly; with an opaque Gaussian in front they differ."""
    def build(logit):
        ps = GaussianSet(
            means=torch.tensor([[0.0219, -0.0219, -3.0], [0.05, 0.02, -4.0], [-0.04, 0.0, -2.0]]),  # centre on a pixel centre
            scales=torch.full((3, 3), math.log(0.2)), quats=torch.tensor([[1.0, 0, 0, 0]] * 3),
            features_dc=torch.tensor([[[0.3, -0.2, 0.9]], [[0.5, 0.5, -0.3]], [[-0.4, 0.8, 0.1]]]),
            features_rest=torch.zeros(3, 15, 3), opacities=torch.tensor([[logit], [1.0], [-1.0]]))
        return Frame(camera=syn.make_camera(64, 48), segments=[Segment(ps, CLS_BACKGROUND)])

    v_img = np.ones((48, 64, 4), np.float32)
    v_alpha = np.ones((48, 64), np.float32)
    res = {}
    for logit in (2.0, 9.0):
        fr = build(logit)
        outs = []
        for cb in (0.999, 0.99):
            o = oracle_c.Oracle(fr, alpha_clamp_fwd=0.999, alpha_clamp_bwd=cb)
            fw = o.forward(class_renders=False)
            g, _ = o.backward(fw, v_img, v_alpha)
            outs.append((fw, g))
        np.testing.assert_array_equal(outs[0][0].img, outs[1][0].img)  # forward never depends on it
        res[logit] = rel_l2(outs[1][1][0]["means"], outs[0][1][0]["means"])
    assert res[2.0] == 0.0
    assert res[9.0] > 1e-3


def test_finite_difference_f64_tiny():
    """float64 central differences on a 5-Gaussian scene validate the autograd oracle itself."""
    fr = syn.make_frame(5, 1, n_per_actor=3, width=32, height=32, seed=5)
    with torch.no_grad():
        fr.segments[0].params.means.copy_(torch.tensor(
            [[0.02, 0.01, -3.0], [-0.03, 0.02, -3.5], [0.01, -0.02, -4.0], [0.0, 0.0, -2.5], [0.03, 0.03, -5.0]]))
        fr.segments[0].params.scales.fill_(math.log(0.02))
        fr.segments[0].params.opacities.fill_(0.3)
        fr.segments[1].params.means.mul_(0.02)
        # the reference detaches means before computing SH view directions (sgn_splatfacto.py:934), so
        # autograd deliberately omits d(colour)/d(mean); zero higher-order SH so finite differences agree
        fr.segments[0].params.features_rest.zero_()
        fr.segments[1].params.features_rest.zero_()
        fr.segments[1].params.scales.fill_(math.log(0.015))
        fr.segments[1].params.opacities.fill_(0.2)
    fr.segments[1].center = np.array([0.01, 0.0, -3.2])
    o = oracle_c.Oracle(fr)
    fw = o.forward()
    assert (fw.radii > 0).all()
    g = torch.Generator().manual_seed(1)
    w_img = torch.rand(32, 32, 4, generator=g).double()
    w_a = torch.rand(32, 32, generator=g).double()

    def loss_of(frame):
        leaves, out = oracle_torch.render(frame, fw.sorted_ids, fw.tile_bins, dtype=torch.float64, use_spec_exp=False)
        return leaves, (out["img"] * w_img).sum() + (out["alpha"] * w_a).sum() + out["object_acc"].sum()

    leaves, L = loss_of(fr)
    L.backward()
    eps = 1e-6
    for seg, key, idx in [(0, "means", (1, 0)), (0, "scales", (2, 1)), (0, "quats", (0, 2)), (0, "opacities", (3, 0)),
                          (0, "features_rest", (1, 4, 2)), (1, "means", (0, 1)), (1, "quats", (1, 3)),
                          (1, "features_dc", (2, 3, 1)), (1, "scales", (0, 0))]:
        base = getattr(fr.segments[seg].params, key)
        analytic = leaves[seg][key].grad[idx].item()
        vals = []
        for sgn in (+1, -1):
            fr2 = Frame(fr.camera, [Segment(s.params.detach_clone(), s.cls, s.rot, s.center, s.idft) for s in fr.segments])
            p = getattr(fr2.segments[seg].params, key).double()
            p[idx] += sgn * eps
            # keep float64 precision: patch after compose() casts -> inject via a float64 tensor
            setattr(fr2.segments[seg].params, key, p)
            _, Lp = loss_of(fr2)
            vals.append(Lp.item())
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - analytic) <= 1e-5 * max(1.0, abs(analytic)), (seg, key, idx, fd, analytic)

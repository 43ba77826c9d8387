"""GPU parity tests: the CUDA path (through the C-ABI of libsgn_raster.so) against the CPU oracle
on identical seeded inputs.

Bars (BASELINE.json north_star): integer / index outputs bit-exact; RGB within 1e-4 max-abs;
gradients within 1e-3 relative L2 per tensor.  Pixels the oracle flags as `fragile` (a skip /
termination decision within a 1e-4 relative margin of flipping under fast-math exp) are excluded
from the max-abs check and bounded separately (DESIGN.md "Parity protocol")."""
import numpy as np
import pytest
import torch

import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200 import raster
from street_gaussians_ns_b200.scene import Frame, Segment
from oracle import oracle_c

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4
GRAD_TOL = 1e-3


def rel_l2(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a)


def to_cuda(frame: Frame, requires_grad=False) -> Frame:
    segs = []
    for s in frame.segments:
        p = s.params.to("cuda")
        if requires_grad:
            p.requires_grad_(True)
        segs.append(Segment(p, s.cls, s.rot, s.center, s.idft, s.name))
    return Frame(frame.camera, segs)


SCENES = {
    "small_actors": dict(n_background=20000, n_actors=6, n_per_actor=1500, width=320, height=240, seed=3,
                         actor_shift=np.array([1.0, 0.0, -1.0])),
    "cfg1_50k_640x480": dict(n_background=50000, n_actors=0, width=640, height=480, seed=0),
    "ragged_edge": dict(n_background=8000, n_actors=2, n_per_actor=500, width=200, height=136, seed=9,
                        actor_shift=np.array([1.5, 0.0, 0.0])),
    # odd image size (no multiple of 16 in either direction), many small actors, rotated camera of the Waymo rig
    "odd_many_actors": dict(n_background=12000, n_actors=9, n_per_actor=300, width=333, height=177, seed=21,
                            actor_shift=np.array([0.5, 0.0, -2.0]), c2w=syn.waymo_rig(4)[6]),
    # dense overlap: long per-tile lists on a small image, so tiles split into 2/4/8 strips and pixels saturate
    "dense_small_image": dict(n_background=60000, n_actors=3, n_per_actor=4000, width=96, height=80, seed=5,
                              actor_shift=np.array([1.0, 0.0, 2.0]), fourier_dim=1),
}


@pytest.fixture(scope="module", params=list(SCENES))
def scene(request):
    fr = syn.make_frame(**SCENES[request.param])
    orc = oracle_c.Oracle(fr)  # gsplat clamps: forward 0.999, backward 0.99
    fw = orc.forward()
    return request.param, fr, orc, fw


def test_project_bit_exact_ints_and_records(scene):
    name, fr, orc, fw = scene
    frc = to_cuda(fr)
    s = raster.RenderSettings()
    cs = raster.camera_struct(frc.camera, s)
    table = raster.SegmentTable(frc, [seg.params.tensors() for seg in frc.segments], torch.device("cuda", 0))
    records, radii, tiles_hit, bbox = raster.project_fwd(table, cs, torch.device("cuda", 0))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(radii.cpu().numpy(), fw.radii)
    np.testing.assert_array_equal(tiles_hit.cpu().numpy(), fw.num_tiles_hit)
    vis = fw.radii > 0
    np.testing.assert_array_equal(bbox.cpu().numpy().astype(np.int32)[vis], fw.tile_bbox[vis])
    rec = records.cpu().numpy()
    # exact section: same IEEE operation sequence on both sides -> identical bits
    np.testing.assert_array_equal(rec[:, 0:2], fw.xys)
    np.testing.assert_array_equal(rec[:, 2:5], fw.conics)
    np.testing.assert_array_equal(rec[:, 9], fw.depths)
    # colour / opacity are only produced for rows the camera sees (nothing reads them otherwise)
    np.testing.assert_allclose(rec[vis, 5], fw.opac[vis], atol=1e-6)
    np.testing.assert_allclose(rec[vis][:, 6:9], fw.rgbs[vis], atol=5e-6)
    assert np.all(rec[~vis][:, 5:9] == 0)
    aux = rec[:, 10].view(np.int32)
    np.testing.assert_array_equal((aux >> 3) & 1, fw.cls)
    np.testing.assert_array_equal((aux >> 4) & 1, vis.astype(np.int32))


def test_binning_exact_order_and_noop_culling(scene):
    """The product's per-tile lists are the oracle's (gsplat AABB) lists, in the same order, minus
    entries that no pixel centre of the tile can accept (alpha < 1/255 everywhere): bit-exact
    subsequence, and every dropped entry is a no-op according to the oracle."""
    name, fr, orc, fw = scene
    frc = to_cuda(fr)
    s = raster.RenderSettings()
    cs = raster.camera_struct(frc.camera, s)
    dev = torch.device("cuda", 0)
    table = raster.SegmentTable(frc, [seg.params.tensors() for seg in frc.segments], dev)
    records, radii, tiles_hit, bbox = raster.project_fwd(table, cs, dev)
    M, sorted_ids, tile_bins = raster.bin_and_sort(cs, records, radii, tiles_hit, bbox)
    ids = sorted_ids.cpu().numpy()[:M]
    tb = tile_bins.cpu().numpy()
    any_valid = orc.entry_any_valid(fw)
    assert M <= fw.M and M >= int(any_valid.sum())
    assert M < 0.9 * fw.M  # the culling actually removes work on these scenes
    kept_total = 0
    for t in range(tb.shape[0]):
        mine = ids[tb[t, 0]: tb[t, 1]]
        ref = fw.sorted_ids[fw.tile_bins[t, 0]: fw.tile_bins[t, 1]]
        av = any_valid[fw.tile_bins[t, 0]: fw.tile_bins[t, 1]]
        # subsequence with identical order: mark which reference entries were kept
        keep = np.zeros(len(ref), bool)
        pos = 0
        for g in (mine & 0x7FFFFFFF):
            while pos < len(ref) and ref[pos] != g:
                pos += 1
            assert pos < len(ref), f"tile {t}: entry {g} not in the reference list (or out of order)"
            keep[pos] = True
            pos += 1
        assert not np.any(av & ~keep), f"tile {t}: dropped an entry some pixel accepts"
        kept_total += keep.sum()
        np.testing.assert_array_equal((mine < 0).astype(np.int32), fw.cls[mine & 0x7FFFFFFF])  # bit 31 = object class
    assert kept_total == M


def _render(frc, training=True, sky=None, **kw):
    s = raster.RenderSettings(training=training, **kw)
    return raster.render_frame(frc, s, sky=sky)


def test_blend_forward_parity(scene):
    name, fr, orc, fw = scene
    frc = to_cuda(fr)
    out, holder = _render(frc)
    torch.cuda.synchronize()
    H, W = fr.camera.height, fr.camera.width
    alpha = (1 - fw.final_T)
    rgb_ref, a_ref, depth_ref = oracle_c.post_ops(torch.from_numpy(fw.img), torch.from_numpy(alpha), None, True)
    ok = fw.fragile == 0
    print(f"[{name}] fragile fraction {1 - ok.mean():.5f}")
    assert ok.mean() >= 0.99  # small images with long lists (dense_small_image) are the worst case; full size: <= 0.4 %
    rgb = out["rgb"].cpu().numpy()
    assert np.abs(rgb - rgb_ref.numpy())[ok].max() <= RGB_TOL
    assert np.abs(out["accumulation"].cpu().numpy()[..., 0] - alpha)[ok].max() <= RGB_TOL
    d, dr = out["depth"].cpu().numpy()[..., 0], depth_ref.numpy()[..., 0]
    sel = ok & (alpha > 2e-3)
    assert (np.abs(d - dr)[sel] / np.maximum(dr[sel], 1.0)).max() <= 1e-3
    assert np.abs(out["object_acc"].cpu().numpy()[..., 0] - (1 - fw.obj_T))[fw.fragile_obj == 0].max() <= RGB_TOL
    assert np.abs(out["background_acc"].cpu().numpy()[..., 0] - (1 - fw.bg_T))[fw.fragile_bg == 0].max() <= RGB_TOL
    # fragile pixels are still sane: one marginal Gaussian (alpha ~ 1/255) more or less
    print(f"[{name}] max |rgb - oracle| over ALL pixels {np.abs(rgb - rgb_ref.numpy()).max():.2e}")
    assert np.abs(rgb - rgb_ref.numpy()).max() <= 0.02
    assert np.isfinite(rgb).all()


@pytest.mark.parametrize("masked", [True, False])
def test_backward_parity(scene, masked):
    """``masked``: cotangents zeroed on the fragile pixels (the gradient the two sides agree to compute); unmasked: every
    pixel contributes, including those whose marginal skip / termination decision the approximate exp may flip."""
    name, fr, orc, fw = scene
    frc = to_cuda(fr, requires_grad=True)
    out, holder = _render(frc)
    H, W = fr.camera.height, fr.camera.width
    g = torch.Generator().manual_seed(7)
    ok = ((fw.fragile == 0) & (fw.fragile_obj == 0) & (fw.fragile_bg == 0)).astype(np.float32)
    okt = torch.from_numpy(ok) if masked else torch.ones(H, W)
    w_rgb = torch.rand(H, W, 3, generator=g) * okt[..., None]
    w_a = torch.rand(H, W, generator=g) * okt
    w_d = 0.05 * torch.rand(H, W, generator=g) * okt
    w_o = torch.rand(H, W, generator=g) * okt
    w_b = torch.rand(H, W, generator=g) * okt
    loss = ((out["rgb"] * w_rgb.cuda()).sum() + (out["accumulation"][..., 0] * w_a.cuda()).sum()
            + (out["depth"][..., 0] * w_d.cuda()).sum() + (out["object_acc"][..., 0] * w_o.cuda()).sum()
            + (out["background_acc"][..., 0] * w_b.cuda()).sum())
    loss.backward()
    torch.cuda.synchronize()
    # oracle: chain the reference post-ops with torch autograd on the oracle's raw outputs, then C backward
    img = torch.from_numpy(fw.img).requires_grad_(True)
    alpha = torch.from_numpy(1 - fw.final_T).requires_grad_(True)
    rgb_ref, a_ref, depth_ref = oracle_c.post_ops(img, alpha, None, True)
    depth_term = torch.where(alpha[..., None] > 1e-3, depth_ref, torch.zeros_like(depth_ref))
    lref = (rgb_ref * w_rgb).sum() + (a_ref[..., 0] * w_a).sum() + (depth_term[..., 0] * w_d).sum()
    lref.backward()
    grads, rastergrads = orc.backward(fw, img.grad.numpy(), alpha.grad.numpy(), w_o.numpy(), w_b.numpy())
    # per-Gaussian raster gradients first (localises failures), then every parameter tensor
    v = holder.v_records.cpu().numpy()
    assert rel_l2(v[:, 0:2], rastergrads["v_xy"]) <= GRAD_TOL
    assert rel_l2(v[:, 2:5], rastergrads["v_conic"]) <= GRAD_TOL
    assert rel_l2(v[:, 5], rastergrads["v_opac"]) <= GRAD_TOL
    assert rel_l2(v[:, 6:9], rastergrads["v_rgb"]) <= GRAD_TOL
    assert rel_l2(v[:, 9], rastergrads["v_depth"]) <= GRAD_TOL
    for si, (seg, gref) in enumerate(zip(frc.segments, grads)):
        for k in ("means", "scales", "quats", "features_dc", "features_rest", "opacities"):
            got = getattr(seg.params, k).grad
            assert got is not None, (si, k)
            ref = gref[k]
            if np.linalg.norm(ref) == 0:
                assert float(got.abs().max()) == 0.0, (si, k)
                continue
            err = rel_l2(got.cpu().numpy(), ref)
            assert err <= GRAD_TOL, (name, si, k, err)


def test_sky_blend_eval_clamp_and_sky_grad():
    fr = syn.make_frame(**SCENES["small_actors"])
    orc = oracle_c.Oracle(fr)
    fw = orc.forward(class_renders=False)
    H, W = fr.camera.height, fr.camera.width
    g = torch.Generator().manual_seed(3)
    sky = torch.rand(H, W, 3, generator=g)
    for training in (True, False):
        frc = to_cuda(fr, requires_grad=training)
        skyc = sky.cuda().requires_grad_(training)
        out, _ = _render(frc, training=training, sky=skyc, class_streams=False)
        alpha = torch.from_numpy(1 - fw.final_T)
        rgb_ref, _, _ = oracle_c.post_ops(torch.from_numpy(fw.img), alpha, sky, training)
        ok = fw.fragile == 0
        assert np.abs(out["rgb"].detach().cpu().numpy() - rgb_ref.numpy())[ok].max() <= RGB_TOL
        if training:
            out["rgb"].sum().backward()
            ref = (1 - alpha)[..., None].expand(H, W, 3).numpy()
            np.testing.assert_allclose(skyc.grad.cpu().numpy(), ref, atol=1e-5)


def test_empty_and_degenerate_inputs():
    # no Gaussian in view: all behind the camera -> M == 0, outputs are the sky / zeros, grads are zeros
    fr = syn.make_frame(n_background=500, n_actors=0, width=64, height=48, seed=1)
    with torch.no_grad():
        fr.segments[0].params.means[:, 2].abs_()  # +z is behind an OpenGL camera
    frc = to_cuda(fr, requires_grad=True)
    out, holder = _render(frc)
    assert holder.M == 0
    assert float(out["accumulation"].detach().abs().max()) == 0.0
    assert float(out["rgb"].detach().abs().max()) == 0.0
    assert float((out["depth"].detach() - 10.0).abs().max()) == 0.0
    (out["rgb"].sum() + out["accumulation"].sum()).backward()
    for t in frc.segments[0].params.tensors():
        assert float(t.grad.abs().max()) == 0.0


def test_error_paths():
    from street_gaussians_ns_b200 import _lib
    fr = syn.make_frame(n_background=100, n_actors=0, width=64, height=48)
    frc = to_cuda(fr)
    with pytest.raises(_lib.SgnError):  # block width outside gsplat's (1,16]
        _render(frc, block_width=32)
    with pytest.raises(_lib.SgnError):  # CPU tensors: no CPU path
        raster.render_frame(fr, raster.RenderSettings())


def test_full_size_properties_cfg3():
    """BASELINE config 3 at full size: size-independent properties (no oracle at this size in the
    GPU suite: it runs in bench.py's cpu_baseline leg).  Sortedness, bin coverage, alpha range,
    determinism of the forward, and linearity of the backward in the cotangent."""
    fr = syn.config_frame(3)
    frc = to_cuda(fr, requires_grad=True)
    s = raster.RenderSettings()
    cs = raster.camera_struct(frc.camera, s)
    dev = torch.device("cuda", 0)
    table = raster.SegmentTable(frc, [seg.params.tensors() for seg in frc.segments], dev)
    records, radii, tiles_hit, bbox = raster.project_fwd(table, cs, dev)
    M, sorted_ids, tile_bins = raster.bin_and_sort(cs, records, radii, tiles_hit, bbox)
    assert 1_000_000 < M <= int(tiles_hit.sum().item())  # exact tile culling only removes no-op entries
    tb = tile_bins.cpu().numpy()
    nonempty = tb[:, 1] > tb[:, 0]
    assert tb[nonempty][0, 0] == 0 and tb[nonempty][-1, 1] == M
    assert np.all(tb[nonempty][1:, 0] == tb[nonempty][:-1, 1])  # bins tile the list
    depth = records[:, 9][(sorted_ids[:M] & 0x7FFFFFFF).long()].cpu().numpy()
    seg_start = np.zeros(M, bool)
    seg_start[tb[nonempty][:, 0]] = True
    assert np.all((np.diff(depth) >= 0) | seg_start[1:])  # depth-sorted inside every tile
    out1, h1 = raster.render_frame(frc, s)
    out2, _ = raster.render_frame(frc, s)
    for k in ("rgb", "accumulation", "depth", "object_acc", "background_acc"):
        assert torch.equal(out1[k], out2[k]), k  # forward is deterministic
    a = out1["accumulation"]
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    for k in ("object_acc", "background_acc"):
        assert float(out1[k].min()) >= 0.0 and float(out1[k].max()) <= 1.0
    w, v = syn.cotangents(cs.height, cs.width)
    w, v = w.cuda(), v.cuda()
    (out1["rgb"] * w).sum().backward(retain_graph=True)
    g1 = frc.segments[0].params.means.grad.clone()
    frc.segments[0].params.means.grad = None
    (out1["rgb"] * (2 * w)).sum().backward()
    g2 = frc.segments[0].params.means.grad
    assert rel_l2(g2.cpu().numpy(), 2 * g1.cpu().numpy()) < 1e-4  # atomics reorder sums: not bit-exact
    assert torch.isfinite(g2).all()


def test_forward_backward_without_autograd_matches_autograd():
    """raster.forward_backward (the C-ABI stages called back to back) == render_frame + backward()."""
    fr = syn.make_frame(**SCENES["small_actors"])
    frc = to_cuda(fr, requires_grad=True)
    s = raster.RenderSettings()
    H, W = fr.camera.height, fr.camera.width
    w, v = syn.cotangents(H, W)
    w, v = w.cuda(), v.cuda()
    out, holder = raster.render_frame(frc, s)
    torch.autograd.backward([out["rgb"], out["accumulation"], out["object_acc"]], [w, v[..., None], 0.1 * v[..., None]])
    ref = torch.cat([t.grad.reshape(-1) for seg in frc.segments for t in seg.params.tensors()])
    out2, h2 = raster.forward_backward(frc, s, {"rgb": w, "accumulation": v, "object_acc": 0.1 * v}, want_param_grads=True)
    got = torch.cat([g.reshape(-1) for g in h2.param_grads])
    for k in ("rgb", "accumulation", "depth", "object_acc", "background_acc"):
        assert torch.equal(out[k].detach().reshape(-1), out2[k].reshape(-1)), k
    assert rel_l2(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5  # float atomics: summation order differs run to run
    assert h2.grad_arena.numel() >= got.numel()


@pytest.mark.parametrize("scene_name", ["small_actors", "ragged_edge"])
def test_execution_variants_agree(scene_name, monkeypatch):
    """The SGN_TUNE_* variants (row skipping, packed f32x2 slot bodies; include/sgn_raster.h) are the same
    computation up to fp32 rounding: termination decisions (final_idx) must be identical on every
    non-fragile pixel, images within 2e-6, gradients within 1e-5 rel-L2 of the scalar kernels."""
    fr = syn.make_frame(**SCENES[scene_name])
    H, W = fr.camera.height, fr.camera.width
    w, v = syn.cotangents(H, W)
    cots = {"rgb": w.cuda(), "accumulation": v.cuda(), "depth": 0.05 * v.cuda(), "object_acc": 0.1 * v.cuda(),
            "background_acc": 0.1 * v.cuda()}
    res = {}
    for tuning in (0, 1, 4 | 8, 1 | 4 | 8 | 16, 4 | 8 | 32):  # 32: lists materialised as staged entries, moved by cp.async.bulk
        monkeypatch.setenv("SGN_TUNING", str(tuning))
        frc = to_cuda(fr, requires_grad=True)
        out, h = raster.forward_backward(frc, raster.RenderSettings(), cots, want_param_grads=True)
        res[tuning] = ({k: out[k].clone() for k in ("rgb", "accumulation", "depth", "object_acc", "background_acc")},
                       torch.cat([g.reshape(-1) for g in h.param_grads]).clone())
    base_out, base_g = res[0]
    for tuning, (o, g) in res.items():
        if tuning == 0:
            continue
        for k, t in o.items():
            diff = (t - base_out[k]).abs().reshape(-1)
            # a different rounding can flip a threshold decision on a handful of pixels: bound their count
            assert (diff > 2e-6).sum().item() <= max(4, diff.numel() // 20000), (tuning, k, diff.max().item())
        assert rel_l2(g.cpu().numpy(), base_g.cpu().numpy()) < 1e-5, tuning


def test_deterministic_mode_is_bit_repeatable_and_close():
    """RenderSettings(deterministic=True) / SGN_DETERMINISTIC=1: per-Gaussian gradients accumulated in 64-bit fixed point (one
    rounding per addend, integer adds) instead of with float atomics -> the whole gradient arena is bit-identical from run
    to run, and agrees with the float-atomic path to its own run-to-run noise."""
    fr = syn.make_frame(**SCENES["dense_small_image"])
    H, W = fr.camera.height, fr.camera.width
    w, v = syn.cotangents(H, W)
    cots = {"rgb": w.cuda(), "accumulation": v.cuda(), "depth": 0.05 * v.cuda(), "object_acc": 0.1 * v.cuda(),
            "background_acc": 0.1 * v.cuda()}
    frc = to_cuda(fr)
    runs = []
    for _ in range(3):
        out, h = raster.forward_backward(frc, raster.RenderSettings(deterministic=True), cots, want_param_grads=True)
        runs.append((h.v_records.clone(), h.grad_arena.clone()))
    for vr, ga in runs[1:]:
        assert torch.equal(vr, runs[0][0]) and torch.equal(ga, runs[0][1])
    out, h = raster.forward_backward(frc, raster.RenderSettings(deterministic=False), cots, want_param_grads=True)
    assert rel_l2(runs[0][1].cpu().numpy(), h.grad_arena.cpu().numpy()) < 1e-5
    # tiny cotangents (a mean-type loss over 2.4 M pixels): the fixed-point grid follows the cotangents' magnitude
    small = {k: t * 1e-7 for k, t in cots.items()}
    o2, h2 = raster.forward_backward(frc, raster.RenderSettings(deterministic=True), small, want_param_grads=True)
    assert rel_l2(h2.grad_arena.cpu().numpy() * 1e7, runs[0][1].cpu().numpy()) < 1e-5


@pytest.mark.parametrize("C", [5, 8, 19])
def test_extra_channels_parity(C):
    """Generic per-Gaussian channels (render_frame(extra=[N,C]); 8 channels per traversal) against the oracle's generic-C
    blend of the same lists: image <= 1e-4 on non-fragile pixels, gradients w.r.t. the channels and -- through the geometry --
    w.r.t. every parameter tensor <= 1e-3 rel-L2 (the extra loss is the only loss here, so parameter gradients come from
    the extra channels alone)."""
    fr = syn.make_frame(**SCENES["small_actors"])
    orc = oracle_c.Oracle(fr)
    fw = orc.forward(class_renders=False)
    H, W = fr.camera.height, fr.camera.width
    g = torch.Generator().manual_seed(31)
    extra = torch.randn(fw.N, C, generator=g)
    pr = dict(xys=fw.xys, conics=fw.conics, opac=fw.opac)
    img_ref, fT, fi, frag = orc.blend(pr, fw.sorted_ids, fw.tile_bins, extra.numpy())
    assert np.array_equal(fi, fw.final_idx)  # same lists, same weights: the termination does not depend on the channels
    frc = to_cuda(fr, requires_grad=True)
    ex = extra.cuda().requires_grad_(True)
    out, holder = raster.render_frame(frc, raster.RenderSettings(class_streams=False), extra=ex)
    ok = fw.fragile == 0
    got = out["extra"].detach().cpu().numpy()
    assert got.shape == (H, W, C)
    scale = max(1.0, float(np.abs(img_ref).max()))
    assert np.abs(got - img_ref)[ok].max() <= RGB_TOL * scale
    w = torch.rand(H, W, C, generator=g) * torch.from_numpy(ok.astype(np.float32))[..., None]
    (out["extra"] * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    # the Python wrapper Oracle.blend_bwd keeps at most 4 colour channels: call the C entry point with C channels directly
    colors = np.ascontiguousarray(extra.numpy(), np.float32)
    import ctypes as Cc
    N = fw.N
    vc = np.zeros((N, C), np.float64)
    vxy, vcon, vop = np.zeros((N, 2)), np.zeros((N, 3)), np.zeros(N)
    rc = orc.L.sgn_oracle_blend_bwd(
        Cc.c_int(W), Cc.c_int(H), Cc.c_int(16), Cc.c_int(C), oracle_c._p(fw.sorted_ids), oracle_c._p(fw.tile_bins), oracle_c._p(fw.xys),
        oracle_c._p(fw.conics), oracle_c._p(colors), oracle_c._p(fw.opac), oracle_c._p(np.zeros(C, np.float32)), Cc.c_float(0.99),
        oracle_c._p(orc.cls), Cc.c_int(-1), oracle_c._p(np.ascontiguousarray(fw.final_T)), oracle_c._p(np.ascontiguousarray(fw.final_idx)),
        oracle_c._p(np.ascontiguousarray(w.numpy(), np.float32)), oracle_c._p(np.zeros((H, W), np.float32)), oracle_c._p(vxy),
        oracle_c._p(vcon), oracle_c._p(vc), oracle_c._p(vop))
    assert rc == 0
    v_extra_ref = vc
    assert rel_l2(ex.grad.cpu().numpy(), v_extra_ref) <= GRAD_TOL
    v = holder.v_records.cpu().numpy()
    assert rel_l2(v[:, 0:2], vxy) <= GRAD_TOL and rel_l2(v[:, 2:5], vcon) <= GRAD_TOL and rel_l2(v[:, 5], vop) <= GRAD_TOL
    grads = orc.project_bwd(fw, vxy.astype(np.float32), np.zeros(N, np.float32), vcon.astype(np.float32), np.zeros((N, 3), np.float32),
                            vop.astype(np.float32))
    for seg, gref in zip(frc.segments, grads):
        for k in ("means", "scales", "quats", "opacities"):
            ref = gref[k]
            if np.linalg.norm(ref) > 0:
                assert rel_l2(getattr(seg.params, k).grad.cpu().numpy(), ref) <= GRAD_TOL, k


def test_async_binning_identical_results_and_overflow_detection():
    """RenderSettings(async_binning=True): no read-back of the intersection count inside a frame (the capacity comes from
    earlier frames, unused slots are padded behind the last tile).  Same lists, same images, same gradients as the exact path;
    a frame that exceeds the capacity is detected with the next frame and the capacity grows."""
    import warnings
    fr = syn.make_frame(**SCENES["small_actors"])
    H, W = fr.camera.height, fr.camera.width
    w, v = syn.cotangents(H, W)
    cots = {"rgb": w.cuda(), "accumulation": v.cuda(), "object_acc": 0.1 * v.cuda()}
    frc = to_cuda(fr)
    raster._ASYNC_STATE.clear()
    ref_out, ref_h = raster.forward_backward(frc, raster.RenderSettings(deterministic=True), cots, want_param_grads=True)
    stats0 = dict(raster.ASYNC_STATS)
    s = raster.RenderSettings(deterministic=True, async_binning=True)
    out1, h1 = raster.forward_backward(frc, s, cots, want_param_grads=True)   # first frame on the device: learns the count (sync)
    assert isinstance(h1.M, int) and h1.M == ref_h.M
    out2, h2 = raster.forward_backward(frc, s, cots, want_param_grads=True)   # no read-back
    assert isinstance(h2.M, raster.LazyCount) and raster.ASYNC_STATS["frames"] == stats0["frames"] + 1
    for k in ("rgb", "accumulation", "depth", "object_acc", "background_acc"):
        assert torch.equal(out2[k], ref_out[k]), k
    assert torch.equal(h2.grad_arena, ref_h.grad_arena)  # deterministic accumulation: bit-identical
    assert int(h2.M) == ref_h.M and h2.M.capacity >= ref_h.M
    # force an overflow: pretend earlier frames were tiny
    torch.cuda.synchronize()
    st = raster._ASYNC_STATE[str(torch.device("cuda", 0))]
    raster._async_poll(st)
    st["max_m"] = 1000
    monkey_granule, raster.ASYNC_GRANULE = raster.ASYNC_GRANULE, 1024
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        out3, h3 = raster.forward_backward(frc, s, cots)     # truncated lists (capacity 2048 < M): finite, flagged later
        torch.cuda.synchronize()
        assert h3.M.raw() == ref_h.M and int(h3.M) == h3.M.capacity < ref_h.M
        assert bool(torch.isfinite(out3["rgb"]).all())
        out4, h4 = raster.forward_backward(frc, s, cots, want_param_grads=True)   # the poll sees the overflow, the capacity follows
    raster.ASYNC_GRANULE = monkey_granule
    assert raster.ASYNC_STATS["overflows"] == stats0["overflows"] + 1 and any("truncated" in str(c.message) for c in caught)
    assert h4.M.capacity >= ref_h.M
    for k in ("rgb", "accumulation", "object_acc"):
        assert torch.equal(out4[k], ref_out[k]), k
    assert int(st["overflow"].item()) == 1

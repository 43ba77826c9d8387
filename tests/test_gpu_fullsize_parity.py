"""GPU parity at the BENCHMARKED sizes: BASELINE.json configs 2, 3 and one frame of config 4, at full size, against the C
oracle (oracle/sgn_oracle.c; ~2 s per frame on the box's host cores) -- projection bits, per-tile lists, all five images,
and every parameter gradient of every sub-model.

Reported per scene (gpurun_out/fullsize_parity_<scene>.json, copied to profiles/ once per round) and asserted:
  * fragile fraction: pixels where a skip / termination decision is within a 1e-4 RELATIVE margin of flipping under the
    blend kernels' ex2.approx (flagged by the oracle).  Bar: <= 0.5 % of the pixels;
  * max |rgb - oracle| on the non-fragile pixels (bar 1e-4, the north-star's) and on the fragile ones (reported; bounded by
    one marginal Gaussian's contribution);
  * gradients: relative L2 per tensor with cotangents masked to the non-fragile pixels (bar 1e-3, the north-star's) AND with
    UNMASKED cotangents (every pixel contributes; same bar -- measured on the B200: <= 7e-5 on all three scenes,
    profiles/r02a_fullsize_parity_*.json).
"""
import json
import os

import numpy as np
import pytest
import torch

import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200 import raster
from oracle import oracle_c
from tests.test_gpu_parity import rel_l2, to_cuda

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4
GRAD_TOL = 1e-3
GRAD_TOL_UNMASKED = 1e-3
FRAGILE_MAX = 0.005

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg4_frame():
    # camera 106 = frame 21, rig camera 1 (yaw +50 deg): background 1.68 M + the actors with a box in that frame
    return syn.WaymoScene().frame(106)


FULL = {
    "cfg2": lambda: syn.config_frame(2),
    "cfg3": lambda: syn.config_frame(3),
    "cfg4_frame": _cfg4_frame,
}


class _Report(dict):
    def flush(self, name):
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"fullsize_parity_{name}.json"), "w") as f:
            json.dump(self, f, indent=1, sort_keys=True)


@pytest.fixture(scope="module", params=list(FULL))
def full(request):
    oracle_c.lib().sgn_oracle_set_threads(max(1, len(os.sched_getaffinity(0))))
    fr = FULL[request.param]()
    orc = oracle_c.Oracle(fr)
    fw = orc.forward()
    rep = _Report(scene=request.param, N=int(fw.N), M_oracle=int(fw.M), N_visible=int((fw.radii > 0).sum()),
                  segments=len(fr.segments), width=fr.camera.width, height=fr.camera.height)
    yield request.param, fr, orc, fw, rep
    rep.flush(request.param)


def _project(frc):
    s = raster.RenderSettings()
    cs = raster.camera_struct(frc.camera, s)
    dev = torch.device("cuda", 0)
    table = raster.SegmentTable(frc, [seg.params.tensors() for seg in frc.segments], dev)
    return cs, raster.project_fwd(table, cs, dev)


def test_fullsize_project_bits(full):
    name, fr, orc, fw, rep = full
    cs, proj = _project(to_cuda(fr))
    rec = proj.records.cpu().numpy()
    np.testing.assert_array_equal(proj.radii.cpu().numpy(), fw.radii)
    np.testing.assert_array_equal(proj.tiles_hit.cpu().numpy(), fw.num_tiles_hit)
    vis = fw.radii > 0
    np.testing.assert_array_equal(proj.bbox.cpu().numpy().astype(np.int32)[vis], fw.tile_bbox[vis])
    np.testing.assert_array_equal(rec[:, 0:2], fw.xys)
    np.testing.assert_array_equal(rec[:, 2:5], fw.conics)
    np.testing.assert_array_equal(rec[:, 9], fw.depths)
    rep["max_abs_opacity_err"] = float(np.abs(rec[vis, 5] - fw.opac[vis]).max())
    rep["max_abs_colour_err"] = float(np.abs(rec[vis][:, 6:9] - fw.rgbs[vis]).max())
    assert rep["max_abs_opacity_err"] <= 1e-6 and rep["max_abs_colour_err"] <= 5e-6


def test_fullsize_lists_are_exact_subsequences(full):
    """Per-tile lists: the oracle's (gsplat's AABB lists, depth order, stable) minus entries no pixel centre can accept;
    vectorised form of test_gpu_parity.test_binning_exact_order_and_noop_culling."""
    name, fr, orc, fw, rep = full
    frc = to_cuda(fr)
    cs, proj = _project(frc)
    M, sorted_ids, tile_bins = raster.bin_and_sort(cs, proj.records, proj.radii, proj=proj)
    ids = sorted_ids.cpu().numpy()[:M]
    tb = tile_bins.cpu().numpy().astype(np.int64)
    tiles = tb.shape[0]
    mine_len, ref_len = tb[:, 1] - tb[:, 0], (fw.tile_bins[:, 1] - fw.tile_bins[:, 0]).astype(np.int64)
    nonempty = mine_len > 0
    assert tb[nonempty][0, 0] == 0 and tb[nonempty][-1, 1] == M and np.all(tb[nonempty][1:, 0] == tb[nonempty][:-1, 1])
    rn = ref_len > 0
    rtb = fw.tile_bins.astype(np.int64)
    assert rtb[rn][0, 0] == 0 and rtb[rn][-1, 1] == fw.M and np.all(rtb[rn][1:, 0] == rtb[rn][:-1, 1])
    my_key = (np.repeat(np.arange(tiles, dtype=np.int64), mine_len) << 32) | (ids & 0x7FFFFFFF).astype(np.int64)
    ref_key = (np.repeat(np.arange(tiles, dtype=np.int64), ref_len) << 32) | fw.sorted_ids.astype(np.int64)
    order = np.argsort(ref_key, kind="stable")
    srt = ref_key[order]
    assert np.all(srt[1:] != srt[:-1])  # a Gaussian is listed once per tile
    where = np.searchsorted(srt, my_key)
    assert where.max() < len(srt) and np.array_equal(srt[where], my_key), "an entry is not in the oracle's list of its tile"
    pos = order[where]  # position of each of my entries in the oracle's global list
    assert np.all(np.diff(pos) > 0), "order differs from the oracle's inside some tile"
    keep = np.zeros(fw.M, bool)
    keep[pos] = True
    any_valid = orc.entry_any_valid(fw).astype(bool)
    assert not np.any(any_valid & ~keep), "dropped an entry some pixel accepts"
    np.testing.assert_array_equal((ids < 0).astype(np.int32), fw.cls[ids & 0x7FFFFFFF])
    rep.update(M=int(M), M_any_valid=int(any_valid.sum()), max_per_tile=int(mine_len.max()), max_per_tile_oracle=int(ref_len.max()))


def test_fullsize_forward_images(full):
    name, fr, orc, fw, rep = full
    frc = to_cuda(fr)
    out, holder = raster.render_frame(frc, raster.RenderSettings())
    torch.cuda.synchronize()
    alpha = 1 - fw.final_T
    rgb_ref, _, depth_ref = oracle_c.post_ops(torch.from_numpy(fw.img), torch.from_numpy(alpha), None, True)
    rgb_ref, depth_ref = rgb_ref.numpy(), depth_ref.numpy()[..., 0]
    ok = fw.fragile == 0
    rgb = out["rgb"].cpu().numpy()
    err = np.abs(rgb - rgb_ref).max(axis=2)
    acc_err = np.abs(out["accumulation"].cpu().numpy()[..., 0] - alpha)
    d = out["depth"].cpu().numpy()[..., 0]
    sel = ok & (alpha > 2e-3)
    depth_rel = np.abs(d - depth_ref)[sel] / np.maximum(depth_ref[sel], 1.0)
    obj_err = np.abs(out["object_acc"].cpu().numpy()[..., 0] - (1 - fw.obj_T))
    bg_err = np.abs(out["background_acc"].cpu().numpy()[..., 0] - (1 - fw.bg_T))
    rep.update(
        fragile_frac=float(1 - ok.mean()), fragile_frac_object=float((fw.fragile_obj != 0).mean()),
        fragile_frac_background=float((fw.fragile_bg != 0).mean()),
        rgb_max_err_nonfragile=float(err[ok].max()), rgb_max_err_fragile=float(err[~ok].max()) if (~ok).any() else 0.0,
        rgb_fragile_pixels_over_tol=int((err[~ok] > RGB_TOL).sum()), rgb_max_err_all=float(err.max()),
        acc_max_err_nonfragile=float(acc_err[ok].max()), acc_max_err_all=float(acc_err.max()),
        depth_max_rel_err_nonfragile=float(depth_rel.max()),
        object_acc_max_err_nonfragile=float(obj_err[fw.fragile_obj == 0].max()),
        background_acc_max_err_nonfragile=float(bg_err[fw.fragile_bg == 0].max()),
        pixels=int(ok.size))
    rep.flush(name)
    assert rep["fragile_frac"] <= FRAGILE_MAX, rep["fragile_frac"]
    assert rep["fragile_frac_object"] <= FRAGILE_MAX and rep["fragile_frac_background"] <= FRAGILE_MAX
    assert rep["rgb_max_err_nonfragile"] <= RGB_TOL
    assert rep["acc_max_err_nonfragile"] <= RGB_TOL
    assert rep["depth_max_rel_err_nonfragile"] <= 1e-3
    assert rep["object_acc_max_err_nonfragile"] <= RGB_TOL and rep["background_acc_max_err_nonfragile"] <= RGB_TOL
    # a fragile pixel differs by at most one marginal Gaussian: alpha ~ 1/255 of a colour <= ~2 (clamped SH), or the tail
    # behind a termination at T ~ 1e-4
    # (measured: 1.2e-3 on 2-5 pixels of 2.46 M; profiles/r02a_fullsize_parity_*.json)
    assert rep["rgb_max_err_all"] <= 5e-3, rep["rgb_max_err_all"]
    assert rep["rgb_fragile_pixels_over_tol"] <= 50
    assert np.isfinite(rgb).all()


def _oracle_grads(orc, fw, w_rgb, w_a, w_d, w_o, w_b):
    img = torch.from_numpy(fw.img).requires_grad_(True)
    alpha = torch.from_numpy(1 - fw.final_T).requires_grad_(True)
    rgb_ref, a_ref, depth_ref = oracle_c.post_ops(img, alpha, None, True)
    depth_term = torch.where(alpha[..., None] > 1e-3, depth_ref, torch.zeros_like(depth_ref))
    lref = (rgb_ref * w_rgb).sum() + (a_ref[..., 0] * w_a).sum() + (depth_term[..., 0] * w_d).sum()
    lref.backward()
    return orc.backward(fw, img.grad.numpy(), alpha.grad.numpy(), w_o.numpy(), w_b.numpy())


@pytest.mark.parametrize("masked", [True, False])
def test_fullsize_all_gradients(full, masked):
    name, fr, orc, fw, rep = full
    frc = to_cuda(fr, requires_grad=True)
    out, holder = raster.render_frame(frc, raster.RenderSettings())
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
    grads, rg = _oracle_grads(orc, fw, w_rgb, w_a, w_d, w_o, w_b)
    v = holder.v_records.cpu().numpy()
    tag = "masked" if masked else "unmasked"
    raster_err = {"v_xy": rel_l2(v[:, 0:2], rg["v_xy"]), "v_conic": rel_l2(v[:, 2:5], rg["v_conic"]),
                  "v_opacity": rel_l2(v[:, 5], rg["v_opac"]), "v_rgb": rel_l2(v[:, 6:9], rg["v_rgb"]),
                  "v_depth": rel_l2(v[:, 9], rg["v_depth"])}
    worst = {}
    per_tensor = []
    for si, (seg, gref) in enumerate(zip(frc.segments, grads)):
        for k in ("means", "scales", "quats", "features_dc", "features_rest", "opacities"):
            got = getattr(seg.params, k).grad
            assert got is not None, (si, k)
            ref = gref[k]
            if np.linalg.norm(ref) == 0:
                assert float(got.abs().max()) == 0.0, (si, k)
                continue
            e = rel_l2(got.cpu().numpy(), ref)
            per_tensor.append((si, k, e))
            worst[k] = max(worst.get(k, 0.0), e)
    rep[f"grad_rel_l2_{tag}"] = {"raster": {k: float(x) for k, x in raster_err.items()},
                                 "worst_per_parameter": {k: float(x) for k, x in worst.items()},
                                 "tensors_compared": len(per_tensor),
                                 "worst_tensor": max(per_tensor, key=lambda t: t[2])[:2] + (float(max(per_tensor, key=lambda t: t[2])[2]),)}
    rep.flush(name)
    tol = GRAD_TOL if masked else GRAD_TOL_UNMASKED
    for k, e in raster_err.items():
        assert e <= tol, (name, tag, k, e)
    for si, k, e in per_tensor:
        assert e <= tol, (name, tag, si, k, e)

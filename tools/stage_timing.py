"""Per-stage CUDA-event timing of the hot path on one config (developer tool, not the bench)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from street_gaussians_ns_b200 import raster  # noqa: E402
from street_gaussians_ns_b200.scene import Frame, Segment  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-class", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    fr = syn.config_frame(a.cfg)
    frc = Frame(fr.camera, [Segment(s.params.to(dev), s.cls, s.rot, s.center, s.idft) for s in fr.segments])
    s = raster.RenderSettings(class_streams=not a.no_class)
    cs = raster.camera_struct(frc.camera, s)
    bo = raster.blend_opts(s, False)
    params = [seg.params.tensors() for seg in frc.segments]
    H, W = cs.height, cs.width
    w, v = syn.cotangents(H, W)
    vd = dict(rgb=w.to(dev), accumulation=v.to(dev), depth=None,
              object_acc=(0.1 * v).to(dev) if s.class_streams else None, background_acc=None)
    names = ["table", "project_fwd", "bin_sort", "blend_fwd", "blend_bwd", "project_bwd"]
    acc = {n: 0.0 for n in names}
    M = 0
    for it in range(a.iters + 3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record()
        table = raster.SegmentTable(frc, params, dev); ev[1].record()
        records, radii, tiles_hit, bbox = raster.project_fwd(table, cs, dev); ev[2].record()
        M, sorted_ids, tile_bins = raster.bin_and_sort(cs, records, radii, tiles_hit, bbox); ev[3].record()
        oi = ob = None
        if s.class_streams:
            oi, ob = raster.class_lists(cs, M, sorted_ids, tile_bins)
        out = raster.blend_fwd(cs, bo, records, sorted_ids, tile_bins, None, oi, ob); ev[4].record()
        v_records, _ = raster.blend_bwd(cs, bo, records, sorted_ids, tile_bins, out, None, vd, False, oi, ob); ev[5].record()
        grads, arena = raster.project_bwd(table, params, cs, records, radii, v_records); ev[6].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1])
    res = {n: round(acc[n] / a.iters, 4) for n in names}
    res["total_ms"] = round(sum(res.values()), 4)
    res.update(N=table.N, M=M, n_visible=int((radii > 0).sum().item()), cfg=a.cfg, class_streams=s.class_streams,
               mean_alpha=float(out["accumulation"].mean().item()))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

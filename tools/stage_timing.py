"""Per-stage CUDA-event timing of the hot path on one config (developer tool, not the bench)."""
import argparse
import json
import os
import sys
import time

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
    ap.add_argument("--bg-grad", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    fr = syn.config_frame(a.cfg)
    frc = Frame(fr.camera, [Segment(s.params.to(dev).requires_grad_(True), s.cls, s.rot, s.center, s.idft, s.name)
                            for s in fr.segments])
    s = raster.RenderSettings(class_streams=not a.no_class)
    H, W = fr.camera.height, fr.camera.width
    w, v = syn.cotangents(H, W)
    w, v = w.to(dev), v.to(dev)[..., None]

    leaves = [t for sg in frc.segments for t in sg.params.tensors()]

    def step():
        out, holder = raster.render_frame(frc, s)
        outs, gr = [out["rgb"], out["accumulation"]], [w, v]
        if s.class_streams:
            outs.append(out["object_acc"]); gr.append(0.1 * v)
            if a.bg_grad:
                outs.append(out["background_acc"]); gr.append(0.1 * v)
        torch.autograd.backward(outs, gr)
        for t in leaves:
            t.grad = None
        return out, holder

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    raster.TIMER = raster.StageTimer()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        out, holder = step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.iters * 1e3
    res = {k: round(val, 4) for k, val in sorted(raster.TIMER.mean_ms().items())}
    res["gpu_sum_ms"] = round(sum(res.values()), 4)
    res["wall_ms_per_step"] = round(wall, 3)
    res.update(N=holder.records.shape[0], M=holder.M, n_visible=int((holder.radii > 0).sum().item()), cfg=a.cfg,
               class_streams=s.class_streams)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

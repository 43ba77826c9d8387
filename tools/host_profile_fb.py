"""Wall-clock + cProfile of raster.forward_backward (the bench's `value` loop)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from street_gaussians_ns_b200 import raster  # noqa: E402
from street_gaussians_ns_b200.scene import Frame, Segment  # noqa: E402

dev = torch.device("cuda", 0)
fr = syn.config_frame(3)
frc = Frame(fr.camera, [Segment(s.params.to(dev).requires_grad_(True), s.cls, s.rot, s.center, s.idft, s.name) for s in fr.segments])
s = raster.RenderSettings()
w, v = syn.cotangents(fr.camera.height, fr.camera.width)
w, v = w.to(dev), v.to(dev)
cot = {"rgb": w, "accumulation": v, "object_acc": 0.1 * v}


def loop(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        raster.forward_backward(frc, s, cot)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


loop(5)
print("no timer  ms/step", loop(30))
raster.TIMER = raster.StageTimer()
print("with timer ms/step", loop(30))
raster.TIMER = None
print("no timer  ms/step", loop(30))
pr = cProfile.Profile()
pr.enable()
loop(30)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

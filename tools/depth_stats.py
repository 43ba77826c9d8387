"""List-length / traversal-depth statistics of one cfg frame (developer tool): what the blend kernels' critical path is."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from street_gaussians_ns_b200 import raster  # noqa: E402
from street_gaussians_ns_b200.scene import Frame, Segment  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
fr = syn.config_frame(cfg)
frc = Frame(fr.camera, [Segment(s.params.to(dev), s.cls, s.rot, s.center, s.idft, s.name) for s in fr.segments])
out, h = raster.forward_backward(frc, raster.RenderSettings(), {})
td = h.tile_depth.cpu().numpy()
bins = h.tile_bins.cpu().numpy()
ln = bins[:, 1] - bins[:, 0]
def stats(name, a):
    a = np.sort(a)[::-1]
    print(f"{name:12s} sum {a.sum():9d} max {a[0]:6d} top8 {a[:8].tolist()} p99 {int(np.percentile(a, 99))} p90 {int(np.percentile(a, 90))} median {int(np.median(a))} nonzero {(a > 0).sum()}")
stats("main len", ln)
if td is not None:
    for k, nm in enumerate(("main depth", "obj depth", "bg depth")):
        stats(nm, td[k])

"""cProfile of the host side of one training step (developer tool)."""
import cProfile
import os
import pstats
import sys

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
w, v = w.to(dev), v.to(dev)[..., None]
leaves = [t for sg in frc.segments for t in sg.params.tensors()]


def step():
    out, holder = raster.render_frame(frc, s)
    torch.autograd.backward([out["rgb"], out["accumulation"], out["object_acc"]], [w, v, 0.1 * v])
    for t in leaves:
        t.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)

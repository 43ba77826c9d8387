"""cProfile of the host side of one e2e training step through the model API (developer tool)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from street_gaussians_ns_b200.model import ActorPose, SceneGraphConfig, SceneGraphRasterModel  # noqa: E402

dev = torch.device("cuda", 0)
fr = syn.config_frame(3)
bg = fr.segments[0].params.to(dev)
actors = {s.name.replace("object_", ""): s.params.to(dev) for s in fr.segments[1:]}
import numpy as np  # noqa: E402
base = [(s.name.replace("object_", ""), np.asarray(s.rot, np.float64), np.asarray(s.center, np.float64)) for s in fr.segments[1:]]
frame_list = list(range(85))


def boxes_at(t):  # a new timestamp every step: fresh boxes with new rotations (bench.py's e2e)
    k = int(t)
    a = 2e-4 * (k % 1000)
    Ry = np.array([[np.cos(a), 0.0, np.sin(a)], [0.0, 1.0, 0.0], [-np.sin(a), 0.0, np.cos(a)]])
    return [ActorPose(n, Ry @ r, c + np.array([0.0, 0.0, -0.01 * (k % 50)]), k % 85, frame_list) for n, r, c in base]


model = SceneGraphRasterModel(bg, actors, SceneGraphConfig(use_sky_sphere=False, ssim_lambda=0.0), poses_at=boxes_at).to(dev)
model.train()
model.step = 30000
H, W = fr.camera.height, fr.camera.width
gt_host = (torch.rand(H, W, 3) * 255).to(torch.uint8).pin_memory()
params = list(model.parameters())
ring = torch.zeros(1).pin_memory()


COUNTER = [0]


def step():
    COUNTER[0] += 1
    gt = gt_host.to(dev, non_blocking=True)
    out = model.get_outputs(syn.make_camera(W, H, c2w=np.asarray(fr.camera.c2w), time=float(COUNTER[0])))
    loss = sum(model.get_loss_dict(out, {"image": gt}).values())
    loss.backward()
    ring[0:1].copy_(loss.detach().reshape(1), non_blocking=True)  # async D2H, as bench.py does
    for p in params:
        p.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(30):
    step()
torch.cuda.synchronize()
print("e2e ms/step", (time.perf_counter() - t0) / 30 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)

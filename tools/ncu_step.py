"""One cfg3 training step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` (developer tool).

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/step python tools/ncu_step.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from street_gaussians_ns_b200 import raster  # noqa: E402
from street_gaussians_ns_b200.optim import FusedAdam  # noqa: E402
from street_gaussians_ns_b200.scene import Frame, Segment  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    fr = syn.config_frame(a.cfg)
    frc = Frame(fr.camera, [Segment(s.params.to(dev), s.cls, s.rot, s.center, s.idft, s.name) for s in fr.segments])
    s = raster.RenderSettings()
    w, v = syn.cotangents(fr.camera.height, fr.camera.width)
    cots = {"rgb": w.to(dev), "accumulation": v.to(dev), "object_acc": 0.1 * v.to(dev)}
    opt = FusedAdam([seg.params.tensors() for seg in frc.segments], lrs={k: 0.0 for k in raster_param_names()})

    def step():
        out, holder = raster.forward_backward(frc, s, cots)
        opt.step(holder.grad_arena)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def raster_param_names():
    from street_gaussians_ns_b200.scene import PARAM_NAMES
    return PARAM_NAMES


if __name__ == "__main__":
    main()

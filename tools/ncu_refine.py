"""One refinement of the cfg3 background sub-model (decide + apply with both Adam moments) bracketed by
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (developer tool; the same workload bench.py times as `refinement`).

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/refine python tools/ncu_refine.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from street_gaussians_ns_b200.optim import FusedAdam  # noqa: E402
from street_gaussians_ns_b200.scene import Frame, Segment  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    fr = syn.config_frame(2)  # the background sub-model alone (1 M Gaussians)
    frc = Frame(fr.camera, [Segment(s.params.to(dev), s.cls, s.rot, s.center, s.idft, s.name) for s in fr.segments])
    adam = FusedAdam([seg.params.tensors() for seg in frc.segments])
    H, W = fr.camera.height, fr.camera.width
    bench.measure_refinement(frc, adam, H, W, dev, reps=2, cpu_baseline=False)  # warm-up (allocator, first launches)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    res = bench.measure_refinement(frc, adam, H, W, dev, reps=1, cpu_baseline=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(res)


if __name__ == "__main__":
    main()

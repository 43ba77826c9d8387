"""Smallest possible GPU check of training.warm_up_refinement (one refinement of a throwaway model)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200.model import SceneGraphConfig, SceneGraphRasterModel
from street_gaussians_ns_b200.training import warm_up_refinement
dev = torch.device("cuda", 0)
sc = syn.WaymoScene(scale=0.05)
model = SceneGraphRasterModel(sc.background.to(dev), {k: v.to(dev) for k, v in sc.actors.items()}, SceneGraphConfig(use_sky_sphere=False, num_train_data=425)).to(dev)
out = []
for i in range(2):
    t0 = time.perf_counter()
    info = warm_up_refinement(model)
    out.append(round((time.perf_counter() - t0) * 1e3, 2))
print(json.dumps({"ms_first_second": out, "rows_before": info["rows_before"], "rows_after": info["rows_after"]}))

"""How many integer decisions of the exact section depend on WHICH exp() evaluates exp(scales)?

The CUDA kernel, the C oracle and the torch oracle share one fixed polynomial (sgn_expf_exact / sgn_expf_spec), the
reference calls torch.exp (sgn_splatfacto.py:857).  Bit-equal radii / num_tiles_hit between kernel and oracle therefore
show kernel == oracle; this tool bounds kernel vs reference: it re-runs the oracle's projection of a full-size config
with libm expf and with the polynomial's value moved one ulp up / down, and counts the Gaussians whose radius,
num_tiles_hit or tile AABB change (CPU only, runs in the build container).

    python tools/exp_sensitivity.py --cfg 3 > profiles/r02a_exp_sensitivity_cfg3.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    args = ap.parse_args()
    import street_gaussians_ns_b200.synthetic as syn
    from oracle import oracle_c
    fr = syn.config_frame(args.cfg)
    orc = oracle_c.Oracle(fr)
    L = oracle_c.lib()
    L.sgn_oracle_set_exp_mode(0)
    base = orc.project()
    vis = base["radii"] > 0
    out = {"config": args.cfg, "N": int(orc.N), "N_visible": int(vis.sum()),
           "sum_num_tiles_hit": int(base["num_tiles_hit"].astype(np.int64).sum()), "modes": {}}
    for mode, name in ((1, "libm_expf"), (2, "spec_plus_1ulp"), (3, "spec_minus_1ulp")):
        L.sgn_oracle_set_exp_mode(mode)
        pr = orc.project()
        d_rad = pr["radii"] != base["radii"]
        d_tiles = pr["num_tiles_hit"] != base["num_tiles_hit"]
        d_bbox = (pr["tile_bbox"] != base["tile_bbox"]).any(axis=1)
        d_vis = (pr["radii"] > 0) != vis
        out["modes"][name] = {
            "radii_changed": int(d_rad.sum()), "num_tiles_hit_changed": int(d_tiles.sum()), "tile_bbox_changed": int(d_bbox.sum()),
            "visibility_changed": int(d_vis.sum()), "max_abs_radius_change": int(np.abs(pr["radii"] - base["radii"]).max()),
            "intersections_delta": int(pr["num_tiles_hit"].astype(np.int64).sum() - base["num_tiles_hit"].astype(np.int64).sum()),
            # relative to the row's largest conic entry (the off-diagonal term can be arbitrarily close to zero)
            "max_conic_change_rel_to_row": float(np.max(np.abs(pr["conics"][vis & ~d_vis] - base["conics"][vis & ~d_vis]).max(axis=1)
                                                        / np.abs(base["conics"][vis & ~d_vis]).max(axis=1))),
            "frac_of_visible": float(d_rad.sum() / max(int(vis.sum()), 1)),
        }
    L.sgn_oracle_set_exp_mode(0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""How much of the alpha-blend work is spent on (pixel, Gaussian) pairs that cannot contribute -- and how much a finer
traversal granularity would save (CPU analysis with the oracle; design input for the blend kernels).

The blend kernels are instruction-bound: their cost is the number of (pixel, entry) pairs evaluated.  Today a warp
evaluates every surviving entry of a 16x16 tile's list for all 256 pixels, until every pixel of the tile has terminated.
This tool replays the forward traversal of one frame per tile and counts

  useful     pairs with alpha >= 1/255 at a pixel that has not terminated yet (what gsplat's per-pixel loop blends),
  tile       pairs evaluated when the unit is the 16x16 tile (exact tile-level culling, tile-level termination) = today,
  <shape>    pairs evaluated when lists, culling and termination are kept per sub-block of the tile
             (8x8 quadrants, 16x4 / 4x16 strips, 8x4, 4x4 ...).

    python tools/subtile_stats.py --cfg 3 --scale 1.0      (a few minutes on 8 cores; --scale 0.25 for a quick look)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import street_gaussians_ns_b200.synthetic as syn  # noqa: E402
from oracle import oracle_c  # noqa: E402

SHAPES = {"16x16": (16, 16), "8x8": (8, 8), "16x4": (4, 16), "4x16": (16, 4), "8x4": (4, 8), "4x4": (4, 4)}  # name: (rows, cols)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--max-tiles", type=int, default=0, help="> 0: a strided sample of that many tiles")
    a = ap.parse_args()
    t0 = time.time()
    fr = syn.config_frame(a.cfg, scale=a.scale)
    if a.scale != 1.0:  # config_frame shrinks the counts; shrink the image too
        fr.camera = syn.make_camera(int(1920 * a.scale) // 16 * 16, int(1280 * a.scale) // 16 * 16, time=fr.camera.time)
    orc = oracle_c.Oracle(fr)
    fw = orc.forward(class_renders=False)
    H, W = fr.camera.height, fr.camera.width
    tx_n, ty_n = (W + 15) // 16, (H + 15) // 16
    print(f"forward: N={fw.N} M={fw.M} {W}x{H} ({time.time() - t0:.0f}s)", file=sys.stderr)
    xy, con, op = fw.xys, fw.conics, fw.opac
    ids_all, bins = fw.sorted_ids, fw.tile_bins
    tiles = np.arange(tx_n * ty_n)
    if a.max_tiles and a.max_tiles < len(tiles):
        tiles = tiles[:: max(1, len(tiles) // a.max_tiles)]
    tot = {k: 0 for k in SHAPES}
    useful = listed = traversed_entries = 0
    py0, px0 = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    for t in tiles:
        b, e = int(bins[t, 0]), int(bins[t, 1])
        if e <= b:
            continue
        ty, tx = divmod(int(t), tx_n)
        ys, xs = ty * 16 + py0, tx * 16 + px0
        inside = ((ys < H) & (xs < W)).reshape(-1)
        # entries this tile actually traverses: up to the deepest final index of its pixels (+1: the entry that terminates)
        fi = fw.final_idx[ty * 16:min(ty * 16 + 16, H), tx * 16:min(tx * 16 + 16, W)]
        depth = int(min(e - b, max(int(fi.max()) - b + 2, 0)))
        listed += e - b
        if depth <= 0:
            continue
        g = ids_all[b:b + depth].astype(np.int64)
        dx = xy[g, 0][:, None] - (xs.reshape(-1) + 0.5)[None, :]
        dy = xy[g, 1][:, None] - (ys.reshape(-1) + 0.5)[None, :]
        sigma = 0.5 * (con[g, 0][:, None] * dx * dx + con[g, 2][:, None] * dy * dy) + con[g, 1][:, None] * dx * dy
        alpha = np.minimum(0.999, op[g][:, None] * np.exp(-sigma))
        valid = (sigma >= 0) & (alpha >= 1.0 / 255.0) & inside[None, :]                     # [depth, 256]
        # per-pixel termination: entries after the pixel's T drops to <= 1e-4 are not blended
        one_minus = np.where(valid, 1.0 - alpha, 1.0)
        T_incl = np.cumprod(one_minus, axis=0)
        alive = np.vstack([np.ones((1, 256), bool), T_incl[:-1] > 1e-4])                     # pixel still running before entry k
        stop = valid & (T_incl <= 1e-4)
        done_before = np.vstack([np.zeros((1, 256), bool), np.cumsum(stop, axis=0)[:-1] > 0])
        running = alive & ~done_before & inside[None, :]
        useful += int((valid & running).sum())
        traversed_entries += depth
        v3 = valid.reshape(depth, 16, 16)
        r3 = running.reshape(depth, 16, 16)
        for name, (rh, cw) in SHAPES.items():
            # a sub-block evaluates entry k for all its pixels if the entry can touch the block (exact culling at block level)
            # and some pixel of the block is still running (block-level termination)
            vb = v3.reshape(depth, 16 // rh, rh, 16 // cw, cw).any(axis=(2, 4))
            rb = r3.reshape(depth, 16 // rh, rh, 16 // cw, cw).any(axis=(2, 4))
            tot[name] += int((vb & rb).sum()) * rh * cw
    out = {"cfg": a.cfg, "scale": a.scale, "tiles_analysed": int(len(tiles)), "entries_listed": int(listed),
           "entries_traversed": int(traversed_entries), "useful_pairs": int(useful),
           "evaluated_pairs": {k: int(v) for k, v in tot.items()},
           "relative_to_today": {k: round(v / max(tot["16x16"], 1), 4) for k, v in tot.items()},
           "efficiency": {k: round(useful / max(v, 1), 4) for k, v in tot.items()},
           "note": "16x16 = today's unit (exact tile culling + tile-level termination); finer blocks need finer lists "
                   "(more entries to sort: an entry is listed once per block it can touch)",
           "seconds": round(time.time() - t0, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Host emulation of the flattened warp-cooperative tile test in csrc/sgn_touch.cuh (count_touched_tiles): the tiles of the 32
lanes' AABBs laid end to end, 32 tested per step by whichever lane, results routed back to the owners' bit masks.  Checks the
index arithmetic (owner search over the exclusive prefix, bit extraction from the ballots) against the per-lane loops."""
import numpy as np


def flattened(areas, widths, ok):
    """ok(lane, tile_index) -> bool.  Mirrors the kernel's loop structure."""
    mine = np.array([a if a <= 32 else 0 for a in areas])
    pre = np.concatenate([[0], np.cumsum(mine)[:-1]])
    total = int(mine.sum())
    masks = np.zeros(32, np.uint64)
    for base in range(0, total, 32):
        bal = 0
        for lane in range(32):
            f = base + lane
            o = 0
            for step in (16, 8, 4, 2, 1):
                cand = o + step
                if pre[cand] <= f:
                    o = cand
            if f < total:
                ti = f - pre[o]
                w = widths[o]
                row = int((np.float32(ti) + np.float32(0.5)) / np.float32(w))
                assert row == ti // w and ti < mine[o]
                if ok(o, ti):
                    bal |= 1 << lane
        for lane in range(32):
            lo, hi = max(pre[lane], base), min(pre[lane] + mine[lane], base + 32)
            if hi > lo:
                bits = (bal >> (lo - base)) & (0xFFFFFFFF if hi - lo >= 32 else (1 << (hi - lo)) - 1)
                masks[lane] |= np.uint64(bits << (lo - pre[lane]))
    return masks


def test_flattened_equals_per_lane_loops():
    rng = np.random.RandomState(0)
    for trial in range(300):
        kind = trial % 4
        widths = rng.randint(1, 9, 32)
        heights = rng.randint(1, 5, 32)
        areas = widths * heights
        if kind == 1:
            areas[rng.rand(32) < 0.5] = 0           # invisible lanes
        if kind == 2:
            areas[rng.rand(32) < 0.2] = 40          # large AABBs: handled by the other path, contribute nothing here
        if kind == 3:
            areas[:] = 0
            areas[rng.randint(0, 32)] = 32
            widths[:] = 32
        table = rng.rand(32, 64) < 0.6
        got = flattened(areas, widths, lambda lane, ti: bool(table[lane, ti]))
        for lane in range(32):
            a = areas[lane] if areas[lane] <= 32 else 0
            want = sum(1 << ti for ti in range(a) if table[lane, ti])
            assert int(got[lane]) == want, (trial, lane)

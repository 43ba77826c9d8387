"""plan_ranges (dp.py): the chunk ranges of the project backward and the arena slices each one writes -- integer logic (CPU)."""
import numpy as np

from street_gaussians_ns_b200 import dp


def _layout(counts, F):
    widths = [[3, 3, 4, 3 * f, 45, 1] for f in F]
    offs, cur = [], 0
    for n, w in zip(counts, widths):
        row = []
        for k in range(6):
            row.append(cur)
            cur += (n * w[k] + 3) // 4 * 4
        offs.append(row)
    return widths, offs, cur


def test_ranges_cover_every_chunk_and_every_arena_float_once():
    counts = [1_000_003, 10_000, 9_999, 1, 130]
    widths, offs, total = _layout(counts, [1, 5, 5, 5, 5])
    for K in (1, 2, 4, 8):
        plan = dp.plan_ranges(counts, offs, widths, K)
        chunks = sum((n + 127) // 128 for n in counts)
        assert plan[0][0] == 0 and plan[-1][1] == chunks
        assert all(a[1] == b[0] for a, b in zip(plan[:-1], plan[1:]))
        cover = np.zeros(total, np.int32)
        for _, _, slices in plan:
            assert len(slices) <= 48
            for sl in slices:
                o, ln = sl[0], sl[1]
                assert o % 4 == 0 and ln % 4 == 0 and o + ln <= total
                cover[o:o + ln] += 1
                if len(sl) == 5:  # row description of a background slice: rows * width floats fit, first row is chunk-aligned
                    w, r0, rows = sl[2:]
                    assert r0 % 128 == 0 and rows * w <= ln < rows * w + 4
        # every float that holds a gradient is exchanged exactly once; padding floats at most once per neighbouring range
        for s, (n, w) in enumerate(zip(counts, widths)):
            for k in range(6):
                assert np.all(cover[offs[s][k]: offs[s][k] + n * w[k]] == 1), (K, s, k)
        assert cover.max() <= 2
        # the rows of a background range are exactly the rows its chunks cover
        for c0, c1, slices in plan[:-1]:
            r0, r1 = c0 * 128, min(counts[0], c1 * 128)
            assert slices[0] == (offs[0][0] + 3 * r0, (3 * (r1 - r0) + 3) // 4 * 4, 3, r0, r1 - r0)


def test_single_segment_frame():
    counts = [777]
    widths, offs, total = _layout(counts, [1])
    plan = dp.plan_ranges(counts, offs, widths, 4)
    assert plan[0][0] == 0 and plan[-1][1] == 7 and all(len(sl) == 6 for _, _, sl in plan)

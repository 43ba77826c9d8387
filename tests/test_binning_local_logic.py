"""Host emulation of the index arithmetic of the experimental per-tile sort (csrc/binning_local.cu), which was written
without GPU access: the bitonic network's (lo, hi, direction) map, the 64-bit key packing `depth_bits << 32 | row << 1 | class`
with the payload unpacking `row | class << 31`, and the chunked class partition.  The loops below are transliterations of the
kernel's loops (one python iteration per thread); the CUDA kernels themselves are compared with the device-wide path in
tests/test_gpu_zz_experimental.py (gated)."""
import numpy as np
import pytest


def bitonic_like_the_kernel(keys: np.ndarray) -> np.ndarray:
    n = len(keys)
    P = 1
    while P < n:
        P <<= 1
    s = np.full(P, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    s[:n] = keys
    k = 2
    while k <= P:
        j = k >> 1
        while j > 0:
            touched = set()
            for i in range(P >> 1):                      # thread i of the step (pairs of one step are disjoint)
                lo = ((i & ~(j - 1)) << 1) | (i & (j - 1))
                hi = lo | j
                assert lo not in touched and hi not in touched and hi < P
                touched.update((lo, hi))
                up = (lo & k) == 0
                a, b = s[lo], s[hi]
                if (a > b) == up:
                    s[lo], s[hi] = b, a
            assert len(touched) == P
            j >>= 1
        k <<= 1
    return s[:n]


def partition_like_the_kernel(low_words, threads):
    n = len(low_words)
    out0, out1, run0, run1 = {}, {}, 0, 0
    for k0 in range(0, n, threads):
        flags = [1 if (k0 + t < n and (low_words[k0 + t] & 1)) else 0 for t in range(threads)]
        pos = np.concatenate([[0], np.cumsum(flags)[:-1]])   # BlockScan::ExclusiveSum
        total = int(sum(flags))
        for t in range(threads):
            k = k0 + t
            if k < n:
                payload = (low_words[k] >> 1) | ((low_words[k] & 1) << 31)
                if flags[t]:
                    out1[run1 + int(pos[t])] = payload
                else:
                    out0[run0 + (t - int(pos[t]))] = payload
        run1 += total
        run0 += min(threads, n - k0) - total
    return [out0[i] for i in range(run0)], [out1[i] for i in range(run1)]


@pytest.mark.parametrize("n", [1, 2, 3, 17, 100, 257, 1000, 1025])
def test_sort_network_and_key_packing(n):
    rng = np.random.default_rng(n)
    depth = rng.integers(1, 2 ** 31, n, dtype=np.uint64)
    depth[: n // 3] = depth[0]                              # equal depths: the row decides (the stable order of the other path)
    row = rng.permutation(2 ** 20)[:n].astype(np.uint64)
    cls = rng.integers(0, 2, n, dtype=np.uint64)
    keys = (depth << np.uint64(32)) | (row << np.uint64(1)) | cls
    out = bitonic_like_the_kernel(keys)
    assert np.array_equal(out, np.sort(keys))
    low = (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    payload = (low >> 1) | ((low & 1) << 31)
    order = np.lexsort((row, depth))
    assert np.array_equal((payload & 0x7FFFFFFF).astype(np.uint64), row[order])
    assert np.array_equal((payload >> 31).astype(np.uint64), cls[order])


@pytest.mark.parametrize("n,threads", [(1, 256), (255, 256), (256, 256), (257, 256), (700, 256), (2049, 1024)])
def test_class_partition_is_stable(n, threads):
    rng = np.random.default_rng(n)
    rows = rng.permutation(10 ** 6)[:n]
    cls = rng.integers(0, 2, n)
    low = [int(x) for x in ((rows << 1) | cls)]
    a0, a1 = partition_like_the_kernel(low, threads)
    assert a0 == [int(r) for r, c in zip(rows, cls) if c == 0]
    assert a1 == [int(r) | (1 << 31) for r, c in zip(rows, cls) if c == 1]

"""Multi-GPU check of csrc/collective.cu (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/test_collective.py
sgn_allreduce_sym over a symmetric arena -- whole arena, slice lists, average -- against dist.all_reduce on the same data, in both
modes (multimem through the multicast address; peer loads / stores); every rank must end with IDENTICAL bits.  Prints timings."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from street_gaussians_ns_b200 import dp  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 82_500_000  # the 330 MB arena of config 3
    res = {"world": world, "floats": n}
    for use_mc in (True, False):
        ex = dp.SymmetricExchange(n, dev, use_multicast=use_mc)
        tag = "multimem" if ex.multicast_ptr else "p2p"
        if use_mc and not ex.multicast_ptr:
            res["multicast"] = "unavailable on this box"
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        src = torch.randn(n, device=dev, generator=g)
        ref = src.clone()
        dist.all_reduce(ref)
        # whole arena, SUM
        ex.arena.copy_(src)
        ex.all_reduce()
        torch.cuda.synchronize()
        err = float((ex.arena - ref).abs().max() / ref.abs().max())
        assert err < 1e-6, (tag, "sum", err)
        # identical bits on every rank
        chk = ex.arena.view(torch.int32).sum(dtype=torch.int64).reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(int(c) == int(allc[0]) for c in allc), (tag, "replicas differ")
        # slice list + average: untouched floats stay as they were
        ex.arena.copy_(src)
        slices = [(0, 4096), (8192, 1_000_000), (2_000_000, 4), (n - 40_000_000, 40_000_000)]
        ex.all_reduce(slices, average=True)
        torch.cuda.synchronize()
        mask = torch.zeros(n, dtype=torch.bool, device=dev)
        for o, ln in slices:
            mask[o:o + ln] = True
        assert float(((ex.arena - ref / world)[mask]).abs().max() / ref.abs().max()) < 1e-6, (tag, "slices")
        assert torch.equal(ex.arena[~mask], src[~mask]), (tag, "touched outside the slices")
        # range by range on the communication stream
        ex.arena.copy_(src)
        bounds = dp.chunk_bounds(n, 4)
        for k, (lo, hi) in enumerate(bounds):
            ex.after_range(k, [(lo, hi - lo)])
        ex.wait_all()
        torch.cuda.synchronize()
        assert float((ex.arena - ref).abs().max() / ref.abs().max()) < 1e-6, (tag, "ranges")
        ex.arena.copy_(src)  # the same with the consumer waiting range by range (what Adam does)
        for k, (lo, hi) in enumerate(bounds):
            ex.after_range(k, [(lo, hi - lo)])
        for k in range(len(bounds)):
            ex.wait_range(k)
        torch.cuda.synchronize()
        assert float((ex.arena - ref).abs().max() / ref.abs().max()) < 1e-6, (tag, "ranges, waited one by one")
        # timing
        def timeit(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        res[tag + "_ms"] = round(timeit(lambda: ex.all_reduce()), 4)
        for ctas in (32, 64, 148):
            res[f"{tag}_ms_{ctas}ctas"] = round(timeit(lambda: ex.all_reduce(max_ctas=ctas)), 4)
        del ex
    # rows nobody saw are skipped: a background of 1 M rows in the arena layout of project_bwd (6 tensors), ~50 % of the rows visible per
    # rank with a large overlap; unseen rows are all-zero on every rank (as project_bwd leaves them)
    rows = 1_000_000
    widths = [3, 3, 4, 3, 45, 1]
    offs, cur = [], 0
    for w in widths:
        offs.append(cur)
        cur += (rows * w + 3) // 4 * 4
    exs = dp.SymmetricExchange(cur, dev, flag_rows=rows)
    g = torch.Generator(device=dev).manual_seed(7)
    common = torch.rand(rows, device=dev, generator=g) < 0.45            # same on every rank (same seed)
    g2 = torch.Generator(device=dev).manual_seed(1000 + rank)
    own = common | (torch.rand(rows, device=dev, generator=g2) < 0.05)
    src = torch.zeros(cur, device=dev)
    for w, o in zip(widths, offs):
        src[o:o + rows * w].view(rows, w).copy_(torch.randn(rows, w, device=dev, generator=g2) * own[:, None])
    ref = src.clone()
    dist.all_reduce(ref)
    slices = [(o, (rows * w + 3) // 4 * 4, w, 0, rows) for w, o in zip(widths, offs)]
    exs.arena.copy_(src)
    exs.publish_visible(own.to(torch.int32))
    exs.all_reduce(slices, skip_unseen=True)
    torch.cuda.synchronize()
    assert float((exs.arena - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), "row skipping changed the sum"
    def prep():
        exs.publish_visible(own.to(torch.int32))
    res["skip_unseen"] = {"rows": rows, "visible_on_some_rank": float(common.float().mean()) + 0.05,
                          "dense_ms": round(timeit(lambda: exs.all_reduce(slices)), 4),
                          "skipping_ms": round(timeit(lambda: (prep(), exs.all_reduce(slices, skip_unseen=True))), 4), "mode": exs.mode}
    del exs
    auto = dp.SymmetricExchange(n, dev)
    res["auto"] = {"mode": auto.mode, **(auto.tuned or {})}
    del auto
    tmp = torch.zeros(n, device=dev)
    res["nccl_ms"] = round(timeit(lambda: dist.all_reduce(tmp)), 4)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

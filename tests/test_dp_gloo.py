"""world_size-2 gloo test (CPU) of the camera-sharded data-parallel host logic: each rank computes the
gradient arena of ITS camera (with the CPU oracle standing in for the kernels), one all-reduce sums the
arenas, and the result equals the single-process sum over both cameras."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import street_gaussians_ns_b200.synthetic as syn
from street_gaussians_ns_b200 import dp


def _arena_for_camera(cam_index: int) -> torch.Tensor:
    from oracle import oracle_c
    c2w = syn.waymo_rig(4)[cam_index * 5]  # yaw-0 camera advanced 0.5 m per frame
    fr = syn.make_frame(3000, 2, n_per_actor=300, width=96, height=64, seed=2, c2w=c2w,
                        actor_shift=np.array([1.0, 0.0, 0.0]))
    orc = oracle_c.Oracle(fr)
    fw = orc.forward(class_renders=False)
    H, W = fr.camera.height, fr.camera.width
    w, v = syn.cotangents(H, W)
    v4 = np.concatenate([w.numpy(), np.zeros((H, W, 1), np.float32)], axis=2)
    grads, _ = orc.backward(fw, v4, v.numpy())
    flat = [torch.from_numpy(g[k].reshape(-1)) for g in grads for k in
            ("means", "scales", "quats", "features_dc", "features_rest", "opacities")]
    return torch.cat(flat)


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam = dp.camera_for_rank(step=0, rank=rank, world_size=world, num_cameras=4)
    arena = _arena_for_camera(cam)
    dp.allreduce_gradients(arena)
    stats = [torch.full((5,), float(rank + 1)), torch.ones(5), torch.full((5,), float(rank))]
    dp.allreduce_densification_stats(*stats)
    if rank == 0:
        q.put((arena.numpy(), [s.numpy() for s in stats]))
    dist.barrier()
    dist.destroy_process_group()


def _refine_worker(rank: int, world: int, port: int, q, patch: bool = True):
    """Every replica saw a different camera, so its densification statistics differ; after the SUM/SUM/MAX
    all-reduce and with equal seeds the replicas must take IDENTICAL split / duplicate / cull decisions."""
    from street_gaussians_ns_b200 import refine
    from tests.host_harness import use_host_backend
    from tests.test_refine import build_model
    if patch:
        use_host_backend(refine)  # CPU stand-in for the two CUDA entry points (tests only; a spawned process of its own)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    model, _ = build_model(seed=0)
    model.step = 3400
    ranks = [rank] if world > 1 else [0, 1]  # the single process plays both replicas' statistics
    for si, sub in enumerate(model.all_models.values()):
        n = sub.num_points
        per_rank = []
        for r in ranks:
            if si > 0 and (si - 1) % 2 != r:
                continue  # actor si-1 was never in view on replica r since the last refinement: no statistics there
            g = torch.Generator().manual_seed(100 + r)
            vis = torch.randint(1, 5, (n,), generator=g).float()
            per_rank.append((torch.rand(n, generator=g) * vis * 2.5e-6, vis, torch.rand(n, generator=g) * 0.2))
        d = sub.__dict__
        if per_rank:
            d["xys_grad_norm"] = sum(p[0] for p in per_rank)
            d["vis_counts"] = sum(p[1] for p in per_rank)
            d["max_2Dsize"] = torch.stack([p[2] for p in per_rank]).max(dim=0).values
        else:
            d["xys_grad_norm"] = d["vis_counts"] = d["max_2Dsize"] = None
    model.refinement_after(None, 3400, generator=torch.Generator().manual_seed(7), sync_stats=True)
    q.put((rank, [sub.gauss_params["means"].detach().numpy().copy() for sub in model.all_models.values()],
           [sub.gauss_params["scales"].detach().numpy().copy() for sub in model.all_models.values()]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_replicas_take_identical_refinement_decisions_world2(monkeypatch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_refine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (m, sc)) for r, m, sc in (q.get(timeout=300) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import queue as _q
    single = _q.Queue()
    from street_gaussians_ns_b200 import refine
    from tests.host_harness import load_refine_harness
    H = load_refine_harness()
    monkeypatch.setattr(refine, "_backend", lambda: H)
    monkeypatch.setattr(refine, "_require_cuda", lambda t, what: None)
    _refine_worker(0, 1, 0, single, patch=False)   # one process with the summed statistics
    _, means, scales = single.get()
    for a, b, c in zip(got[0][0], got[1][0], means):
        assert a.shape == b.shape == c.shape and a.shape[0] > 0
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    for a, b, c in zip(got[0][1], got[1][1], scales):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)


def _cpu_adam_launch(opt):
    """Test stand-in for FusedAdam.launch: applies the rows of a table with the kernel's arithmetic (adam.cu) to CPU
    tensors, addressing parameters through the table's raw pointers exactly as the kernel does."""
    import ctypes

    def launch(tab, grad_arena):
        g_all, m_all, v_all = grad_arena.numpy(), opt.exp_avg.numpy(), opt.exp_avg_sq.numpy()
        for r in tab:
            n = int(r["numel"])
            p = np.ctypeslib.as_array((ctypes.c_float * n).from_address(int(r["param"])))
            g = g_all[int(r["grad_offset"]):int(r["grad_offset"]) + n]
            m = m_all[int(r["arena_offset"]):int(r["arena_offset"]) + n]
            v = v_all[int(r["arena_offset"]):int(r["arena_offset"]) + n]
            m += r["one_minus_beta1"] * (g - m)
            v *= r["beta2"]
            v += r["one_minus_beta2"] * g * g
            p -= r["step_size"] * (m / (np.sqrt(v) / r["sqrt_bc2"] + r["eps"]))
    return launch


def _pipeline_worker(rank: int, world: int, port: int, q):
    from street_gaussians_ns_b200.optim import FusedAdam
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = lambda n, F: [(n, 3), (n, 3), (n, 4), (n, F, 3), (n, 15, 3), (n, 1)]  # noqa: E731
    results = []
    for mode in ("serial", "pipelined"):
        g = torch.Generator().manual_seed(1)
        params = [[torch.randn(s, generator=g) for s in shapes(n, F)] for n, F in ((301, 1), (57, 5), (40, 5))]
        opt = FusedAdam(params, chunk_elems=256)
        opt.launch = _cpu_adam_launch(opt)
        gr = torch.Generator().manual_seed(10 + rank)   # every replica has its own gradients
        for it in range(3):
            arena = torch.randn(opt.arena_elems, generator=gr)
            present = None if it == 1 else [0, 2]
            if mode == "serial":
                dp.allreduce_gradients(arena, average=True)
                opt.step(arena, present=present, full_layout=True)
            else:
                dp.allreduce_and_step(arena, opt, present, chunks=3)
        results.append(([t.clone() for ps in params for t in ps], opt.exp_avg.clone(), list(opt.steps)))
    (pa, ma, sa), (pb, mb, sb) = results
    ok = all(torch.equal(a, b) for a, b in zip(pa, pb)) and torch.equal(ma, mb) and sa == sb
    q.put((rank, ok, sa, float(sum(t.abs().sum() for t in pa))))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_allreduce_and_adam_equals_serial_world2():
    """dp.allreduce_and_step (all-reduce of range k+1 in flight while Adam runs on range k) gives bit-identical parameters,
    moments and step counts to all-reduce-then-step, on both replicas; tensors are cut at range ends."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in got)
    assert got[0][2] == got[1][2] == [3] * 6 + [1] * 6 + [3] * 6     # sub-model 1 only had a gradient in step 1
    assert got[0][3] == got[1][3] and got[0][3] > 0                   # replicas end up with identical parameters


def test_chunk_bounds_cover_the_arena():
    for total, chunks in ((10_000_000, 4), (4096, 4), (1000, 8), (8, 3)):
        b = dp.chunk_bounds(total, chunks)
        assert b[0][0] == 0 and b[-1][1] == total and all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
        assert all(lo % 4 == 0 for lo, _ in b) and all(hi > lo for lo, hi in b)


def test_camera_assignment():
    seen = [dp.camera_for_rank(s, r, 4, 425) for s in range(3) for r in range(4)]
    assert seen == list(range(12))  # distinct cameras within and across steps
    assert dp.camera_for_rank(200, 3, 8, 425) == (200 * 8 + 3) % 425


def test_allreduce_of_gradient_arena_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    arena, stats = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _arena_for_camera(0) + _arena_for_camera(1)
    assert np.linalg.norm(ref.numpy()) > 0
    np.testing.assert_allclose(arena, ref.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(stats[0], np.full(5, 3.0))   # SUM
    np.testing.assert_array_equal(stats[1], np.full(5, 2.0))   # SUM
    np.testing.assert_array_equal(stats[2], np.full(5, 1.0))   # MAX


def test_single_process_is_a_noop():
    a = torch.arange(4.0)
    assert dp.allreduce_gradients(a) is a

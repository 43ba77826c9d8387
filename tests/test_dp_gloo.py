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

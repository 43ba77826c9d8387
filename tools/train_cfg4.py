"""BASELINE.json configs 4 and 5 as a runnable loop: Waymo-shape synthetic scene (5 cameras x 85 frames, 1.68 M
background + 32 x 10 k actor Gaussians = 2 M), one training step = render one camera + L1 loss against a seeded target
+ Adam (+ densification statistics every step, refinement every ``--refine-every`` steps); with torchrun and N ranks
(config 5) rank r renders camera (step * N + r) mod 425, the gradient arena is all-reduced and averaged, every rank
applies the same update.

    python tools/train_cfg4.py --steps 50                                                   # config 4, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \\
        --master-port 29511 tools/train_cfg4.py --steps 50                                  # config 5

Prints one JSON line (rank 0): training steps/s over all ranks (weak scaling: one camera per rank per step), the
device time per step (CUDA events, max over ranks), Gaussian counts before / after, the losses of the first and last
step.  bench.py's headline stays config 3 (the metric BASELINE.json quotes); this is the tool for configs 4 / 5.
Actors are boxes parked on the road grid of SURVEY.md 8d; an actor has a box in a frame only while the ego vehicle is
within ``--actor-range`` metres of it, so the set of sub-models in view changes from frame to frame."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks Gaussian counts and the image (smoke runs)")
    ap.add_argument("--refine-every", type=int, default=100)
    ap.add_argument("--start-step", type=int, default=600, help="past warmup_length so that refinement is live")
    ap.add_argument("--actor-range", type=float, default=45.0)
    ap.add_argument("--pipeline-chunks", type=int, default=0, help="> 0: all-reduce and Adam pipelined over that many arena ranges")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import street_gaussians_ns_b200.synthetic as syn
    from street_gaussians_ns_b200 import dp
    from street_gaussians_ns_b200.model import ActorPose, SceneGraphConfig, SceneGraphRasterModel
    from street_gaussians_ns_b200.optim import FusedAdam
    from street_gaussians_ns_b200.refine import RefineSettings
    from street_gaussians_ns_b200.training import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "needs a CUDA device (no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)  # the split samples of a refinement come from the global CUDA generator: equal on every replica

    n_bg = int(1_680_000 * args.scale)
    n_act = max(1, int(10_000 * args.scale))
    W, H = max(64, int(1920 * args.scale) // 16 * 16), max(48, int(1280 * args.scale) // 16 * 16)
    num_frames = 85
    bg = syn.make_background(n_bg, seed=0, box=((-40.0, 40.0), (-6.0, 14.0), (-135.0, 45.0))).to(dev)
    actors = {str(a): syn.make_actor(n_act, seed=1 + a).to(dev) for a in range(32)}
    boxes = [syn.actor_pose(a) for a in range(32)]
    rig = syn.waymo_rig(num_frames)  # index = frame * 5 + camera
    cams = [syn.make_camera(W, H, c2w=rig[i], time=float(i // 5)) for i in range(len(rig))]

    pose_cache = {}

    def poses_at(t):
        f = int(t)
        hit = pose_cache.get(f)
        if hit is None:
            ego_z = -0.5 * f
            hit = pose_cache[f] = [ActorPose(str(a), rot, center, f, list(range(num_frames)))
                                   for a, (rot, center) in enumerate(boxes) if abs(center[2] - ego_z) <= args.actor_range]
        return hit

    rs = RefineSettings(refine_every=args.refine_every)
    cfg = SceneGraphConfig(use_sky_sphere=False, ssim_lambda=0.0, full_gradient_arena=world > 1, refine=rs,
                           object_refine=RefineSettings(refine_every=args.refine_every, cull_alpha_thresh=0.005),
                           num_train_data=len(cams), refine_record=True)
    model = SceneGraphRasterModel(bg, actors, cfg, poses_at=poses_at).to(dev)
    model.train()
    opt = FusedAdam(model.optimizer_params())
    step_fn = TrainStep(model, opt, refine_every=args.refine_every, pipeline_chunks=args.pipeline_chunks)
    g = torch.Generator().manual_seed(5)
    gt = (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8).to(dev)  # get_loss_dict consumes uint8 directly
    counts0 = [sub.num_points for sub in model.all_models.values()]

    def one(i):
        step = args.start_step + i
        mine = [cams[dp.camera_for_rank(step, r, world, len(cams))] for r in range(world)]
        return step_fn(step, mine[rank], {"image": gt}, all_cameras=mine if world > 1 else None)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first = None
    for i in range(args.warmup):
        losses = one(i)
        first = first if first is not None else float(sum(v.detach() for v in losses.values()))
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        losses = one(args.warmup + i)
    e1.record()
    sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    dev_ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    last = float(sum(v.detach() for v in losses.values()))
    if rank == 0:
        counts1 = [sub.num_points for sub in model.all_models.values()]
        print(json.dumps({
            "metric": "training steps/s (render 1 camera + L1 + Adam + densification statistics, refinement every "
                      f"{args.refine_every} steps)", "value": world / (float(dev_ms.item()) * 1e-3), "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(dev_ms.item()),
            "wall_ms_per_step": wall_ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg4/5: 5 cameras x 85 frames, 1.68M background + 32 x 10k actor Gaussians, "
                                   f"{W}x{H}, scale {args.scale}", "parallelism": f"camera-sharded dp{world}",
                       "collective": "all-reduce(SUM)/N of the gradient arena (layout of all sub-models)" if world > 1 else "none"},
            "gaussians_before": int(sum(counts0)), "gaussians_after": int(sum(counts1)),
            "submodels_changed": int(sum(a != b for a, b in zip(counts0, counts1))),
            "loss_first": first, "loss_last": last}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

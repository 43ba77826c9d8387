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


def run(steps: int = 50, warmup: int = 5, scale: float = 1.0, refine_every: int = 100, start_step: int = 600,
        actor_range: float = 45.0, pipeline_chunks: int = 0, overlap: bool = False, resident_table: bool = True,
        async_binning: bool = True) -> dict:
    """One measurement.  torch.distributed must already be initialised when WORLD_SIZE > 1.  Returns the result dict on
    rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    import street_gaussians_ns_b200.synthetic as syn
    from street_gaussians_ns_b200 import dp
    from street_gaussians_ns_b200.model import ActorPose, SceneGraphConfig, SceneGraphRasterModel
    from street_gaussians_ns_b200.optim import FusedAdam
    from street_gaussians_ns_b200.refine import RefineSettings
    from street_gaussians_ns_b200.training import TrainStep

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)

    sc = syn.WaymoScene(scale=scale, actor_range=actor_range)
    W, H, num_frames, cams = sc.width, sc.height, sc.num_frames, sc.cameras
    frame_list = list(range(num_frames))

    def poses_at(t):  # fresh objects per call, as a data manager hands them out
        f = int(t)
        return [ActorPose(str(a), rot, center, f, frame_list) for a, rot, center in sc.boxes_at(f)]

    rs = RefineSettings(refine_every=refine_every)
    cfg = SceneGraphConfig(use_sky_sphere=False, ssim_lambda=0.0, full_gradient_arena=world > 1, refine=rs, async_binning=async_binning,
                           object_refine=RefineSettings(refine_every=refine_every, cull_alpha_thresh=0.005),
                           num_train_data=len(cams), refine_record=True)
    model = SceneGraphRasterModel(sc.background.to(dev), {k: v.to(dev) for k, v in sc.actors.items()}, cfg, poses_at=poses_at).to(dev)
    model.train()
    opt = FusedAdam(model.optimizer_params(), reserve_spare=True)  # no cudaMalloc of moment arenas inside the training loop
    step_fn = TrainStep(model, opt, refine_every=refine_every, pipeline_chunks=pipeline_chunks, overlap=overlap)
    g = torch.Generator().manual_seed(5)
    gt = (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8).to(dev)  # get_loss_dict consumes uint8 directly
    counts0 = [sub.num_points for sub in model.all_models.values()]
    times = [float(f) for f in range(num_frames)]
    if resident_table:
        model.prepare_frames(times)
    if world > 1:
        # NCCL sets up its channels lazily, at the first collective of a size class: the statistics exchange of the first
        # refinement would otherwise pay that one-time cost (tens of ms) inside the timed steps
        warm = torch.zeros(1 << 22, device=dev)
        dist.all_reduce(warm, op=dist.ReduceOp.SUM)
        dist.all_reduce(warm, op=dist.ReduceOp.MAX)
        small = torch.zeros(64, device=dev, dtype=torch.int64)
        dist.all_reduce(small, op=dist.ReduceOp.MIN)
        dist.all_reduce(small, op=dist.ReduceOp.MAX)
        del warm, small
    # CUDA loads a kernel's module at its first launch: the first refinement of a process would pay for ~20 kernels nothing else
    # on the step uses (tens of ms of host time).  One refinement of a throwaway model loads them; best effort, process-local
    try:
        from street_gaussians_ns_b200.training import warm_up_refinement
        info = warm_up_refinement(model)
        refine_warm = {"ok": True, "rows_before": info["rows_before"], "rows_after": info["rows_after"]}
    except Exception as e:  # the measurement is still valid without it (the first refinement then includes the module loads)
        refine_warm = {"ok": False, "error": f"{type(e).__name__}: {e}"[:200]}
    if async_binning:
        # without the per-frame read-back of the intersection count the list buffers have a capacity learnt from earlier frames:
        # look at every rig camera at three points of the drive once (no gradients) so that the capacity covers the widest view
        from street_gaussians_ns_b200 import raster
        model.config.async_binning = False
        with torch.no_grad():
            for f in (0, num_frames // 2, num_frames - 1):
                for c in range(5):
                    model.get_outputs(cams[f * 5 + c])
                    raster.async_learn(dev, int(model._holder.M))
        model.config.async_binning = True
        torch.cuda.synchronize()

    def one(i):
        step = start_step + i
        mine = [cams[dp.camera_for_rank(step, r, world, len(cams))] for r in range(world)]
        # after a refinement replaced parameter tensors the model drops the resident table; a timestamp's rows are then
        # re-staged (and kept on the device) the first time it is rendered again
        return step_fn(step, mine[rank], {"image": gt}, all_cameras=mine if world > 1 else None)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first = None
    for i in range(warmup):
        losses = one(i)
        first = first if first is not None else float(sum(v.detach() for v in losses.values()))
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]  # per-step device time: where a refinement's cost lands
    host_ms = []
    profile_refine = os.environ.get("SGN_PROFILE_REFINE") == "1" and rank == 0
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        h0 = time.perf_counter()
        if profile_refine and refine_every > 0 and (start_step + warmup + i) % refine_every == 0:
            import cProfile
            import pstats
            prof = cProfile.Profile()
            prof.enable()
            losses = one(warmup + i)
            torch.cuda.synchronize()
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
        else:
            losses = one(warmup + i)
        marks[i].record()
        host_ms.append((time.perf_counter() - h0) * 1e3)
    e1.record()
    sync()
    step_ms = [round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 3) for i in range(steps)]
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    dev_ms = torch.tensor([e0.elapsed_time(e1) / steps, wall_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    last = float(sum(v.detach() for v in losses.values()))
    replicas_identical = None
    if world > 1:
        try:  # every replica must hold bit-identical parameters after the run: same exchange result, same Adam, same refinements
            chk = torch.stack([p.detach().double().sum() for p in model.all_models["background"].gauss_params.values()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            replicas_identical = bool(torch.equal(lo, hi))
        except Exception:
            replicas_identical = None
    if rank != 0:
        return None
    counts1 = [sub.num_points for sub in model.all_models.values()]
    ms = float(dev_ms[0].item())
    return {
        "metric": "training steps/s (render 1 camera per rank + L1 + backward + gradient all-reduce + fused Adam + densification "
                  f"statistics, refinement every {refine_every} steps; timed steps {start_step + warmup}..{start_step + warmup + steps - 1})", "value": world / (ms * 1e-3), "unit": "steps/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "wall_ms_per_step": float(dev_ms[1].item()),
        "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg{4 if world == 1 else 5}: 5 cameras x 85 frames, {sc.n_bg} background + 32 x {sc.n_act} actor Gaussians, "
                               f"{W}x{H}; rank r renders camera (step*g + r) mod 425; actors have a box within {actor_range} m of the ego vehicle",
                   "parallelism": f"camera-sharded dp{world}", "start_step": start_step, "refine_every": refine_every,
                   "refinement_kernels_loaded_before_timing": refine_warm,
                   "binning": "no host read-back of the intersection count" if async_binning else "one read-back per frame",
                   "segment_table": "device-resident (staged up front; re-staged per timestamp on first use after a refinement)" if resident_table else "host build per frame",
                   "collective": ("all-reduce(AVG) of the gradient arena (layout of all sub-models), "
                                  + ("overlapped with project_bwd / Adam over arena ranges" if overlap else
                                     (f"pipelined with Adam over {pipeline_chunks} ranges" if pipeline_chunks else "serial"))) if world > 1 else "none"},
        "gaussians_before": int(sum(counts0)), "gaussians_after": int(sum(counts1)),
        "submodels_changed": int(sum(a != b for a, b in zip(counts0, counts1))),
        "loss_first": first, "loss_last": last, "replicas_identical": replicas_identical,
        "step_ms_rank0": step_ms, "host_ms_rank0": [round(x, 3) for x in host_ms]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks Gaussian counts and the image (smoke runs)")
    ap.add_argument("--refine-every", type=int, default=100)
    ap.add_argument("--start-step", type=int, default=600, help="past warmup_length so that refinement is live")
    ap.add_argument("--actor-range", type=float, default=45.0)
    ap.add_argument("--pipeline-chunks", type=int, default=0, help="> 0: all-reduce and Adam pipelined over that many arena ranges")
    ap.add_argument("--overlap", action="store_true", help="all-reduce launched per arena range from project_bwd's ranges (dp.OverlappedStep)")
    ap.add_argument("--host-table", action="store_true", help="build the segment table on the host per frame instead of prepare_frames")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "needs a CUDA device (no CPU path)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(args.steps, args.warmup, args.scale, args.refine_every, args.start_step, args.actor_range, args.pipeline_chunks,
              args.overlap, not args.host_table)
    if res is not None:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 300 $TR --master-port 29572 tools/train_cfg4.py --steps 40 --warmup 10 --refine-every 100 --start-step 601 > gpurun_out/train_cfg5_re100.json 2> gpurun_out/train_cfg5_re100.err
timeout 300 python tools/train_cfg4.py --steps 40 --warmup 10 --refine-every 100 --start-step 601 > gpurun_out/train_cfg4_re100.json 2> gpurun_out/train_cfg4_re100.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max")}, "e2e", round(d["e2e"]["value"],1))
print("cfg5", json.dumps(d.get("training_step_cfg5"))[:600])
for f in ("train_cfg5_re100","train_cfg4_re100"):
    try:
        t=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, t["value"], t["ms_per_step"], t["loss_first"], t["loss_last"])
    except Exception as e: print(f, e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY

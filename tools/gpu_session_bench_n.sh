set -u
N=$1
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max")}, "e2e", round(d["e2e"]["value"],1))
print(json.dumps(d.get("collective"))[:500]); print("cfg5", json.dumps(d.get("training_step_cfg5"))[:300])
PY

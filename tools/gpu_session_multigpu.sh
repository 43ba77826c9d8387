set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
T 300 $TR --master-port 29561 tools/test_collective.py > gpurun_out/collective_${N}gpu.json 2> gpurun_out/collective_${N}gpu.err
T 500 $TR --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
T 300 $TR --master-port 29563 bench.py --gpus $N --steps 20 --warmup 5 --no-cfg45 > gpurun_out/bench_${N}gpu_b.json 2> gpurun_out/bench_${N}gpu_b.err
tail -c 1300 gpurun_out/collective_${N}gpu.json; echo; grep -v "NCCL INFO\|^\*\|OMP_NUM" gpurun_out/collective_${N}gpu.err | tail -8 | cut -c1-300
for f in ${N}gpu ${N}gpu_b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max")}, "e2e", round(d["e2e"]["value"],1))
    print("    ", json.dumps(d.get("collective", d["config"].get("collective")))[:900])
    if d.get("training_step_cfg5"): print("    cfg5", json.dumps(d.get("training_step_cfg5"))[:400])
except Exception as e: print("$f", e); print(open("gpurun_out/bench_$f.err").read()[-3000:])
PY
done
cat gpurun_out/session.log

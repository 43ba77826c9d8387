set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
T 300 $TR --master-port 29551 tools/test_collective.py > gpurun_out/collective.json 2> gpurun_out/collective.err
T 400 $TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
SGN_DP_SKIP_UNSEEN=0 T 400 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg45 > gpurun_out/bench_2gpu_noskip.json 2> gpurun_out/bench_2gpu_noskip.err
T 400 $TR --master-port 29554 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg45 > gpurun_out/bench_2gpu_b.json 2> gpurun_out/bench_2gpu_b.err
tail -c 1200 gpurun_out/collective.json; echo; grep -v "NCCL INFO\|^\*\|OMP_NUM" gpurun_out/collective.err | tail -12 | cut -c1-300
for f in 2gpu 2gpu_noskip 2gpu_b; do python - <<PY
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

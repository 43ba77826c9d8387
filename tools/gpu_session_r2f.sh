set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
NCCL_DEBUG=INFO T 300 $TR --master-port 29541 tools/test_collective.py > gpurun_out/collective.json 2> gpurun_out/collective.err
T 400 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_sym.json 2> gpurun_out/bench_2gpu_sym.err
SGN_DP_EXCHANGE=nccl T 400 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg45 > gpurun_out/bench_2gpu_nccl.json 2> gpurun_out/bench_2gpu_nccl.err
SGN_DP_MULTICAST=0 T 400 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg45 > gpurun_out/bench_2gpu_p2p.json 2> gpurun_out/bench_2gpu_p2p.err
CUDA_VISIBLE_DEVICES=0 T 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
tail -c 1500 gpurun_out/collective.json; tail -5 gpurun_out/collective.err | cut -c1-300; grep -i "nvls" gpurun_out/collective.err | head -3
for f in sym nccl p2p; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_2gpu_$f.json"))
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max")}, d["e2e"]["value"], json.dumps(d["config"]["collective"])[:600])
    print("    cfg5", json.dumps(d.get("training_step_cfg5"))[:500])
except Exception as e: print("$f", e); print(open("gpurun_out/bench_2gpu_$f.err").read()[-2500:])
PY
done
tail -4 gpurun_out/gpu_tests.log; cat gpurun_out/session.log

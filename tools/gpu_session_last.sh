# last 2-GPU session of the round: the tests of the code touched last (one GPU), smoke, then bench.py --gpus 2 (cfg5 over one refinement period)
set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 200 python -m pytest tests/test_gpu_adam.py tests/test_gpu_zz_refine.py tests/test_gpu_model.py tests/test_reference_vectors.py -m gpu -q -x > gpurun_out/gpu_tests_last.log 2>&1
T 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
T 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 2 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/gpu_tests_last.log; tail -1 gpurun_out/smoke.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "e2e", round(d["e2e"]["value"],1))
c=d["training_step_cfg5"]; print({k:c.get(k) for k in ("value","ms_per_step","gaussians_before","gaussians_after","replicas_identical")}); print(c.get("step_ms_rank0"))
PY
cat gpurun_out/session.log

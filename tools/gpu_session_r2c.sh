set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 600 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing.log 2>&1
T 100 python tools/stage_timing.py --cfg 2 --iters 20 > gpurun_out/stage_timing_cfg2.log 2>&1
T 200 python tools/host_profile_e2e.py > gpurun_out/host_profile_e2e.log 2>&1
T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -5 gpurun_out/gpu_tests.log; tail -1 gpurun_out/stage_timing.log; tail -1 gpurun_out/stage_timing_cfg2.log; head -50 gpurun_out/host_profile_e2e.log
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_a.json"))
    print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax")}, d["e2e"]["value"], d["e2e"]["resident_table"]["value"])
    print("   cfg4", json.dumps(d.get("training_step_cfg4"))[:400])
except Exception as e: print(e); print(open("gpurun_out/bench_a.err").read()[-1500:])
PY
cat gpurun_out/session.log

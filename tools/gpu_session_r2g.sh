set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
for b in a b c d; do
T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_$b.json 2> gpurun_out/bench_$b.err
done
SGN_ASYNC_BIN=0 T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_sync.json 2> gpurun_out/bench_sync.err
T 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -4 gpurun_out/gpu_tests.log
for b in a b c d sync full; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$b.json").read().strip().splitlines()[-1])
    print("$b", {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax")}, d["e2e"]["value"], d["e2e"]["resident_table"]["value"], d["config"].get("binning","")[:60])
    if d.get("training_step_cfg4"): print("   cfg4", json.dumps(d["training_step_cfg4"])[:300])
except Exception as e: print("$b", e); print(open("gpurun_out/bench_$b.err").read()[-1500:])
PY
done
cat gpurun_out/session.log

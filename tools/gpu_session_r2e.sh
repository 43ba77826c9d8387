set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
T 300 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -x -k "project_bits or subsequences" > gpurun_out/gpu_tests_fullsize_bits.log 2>&1
T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing.log 2>&1
SGN_TUNING=44 T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing_tma.log 2>&1
SGN_DETERMINISTIC=1 T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing_det.log 2>&1
for b in a b c; do
T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_$b.json 2> gpurun_out/bench_$b.err
done
T 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/ncu_step.py > /dev/null 2>&1
tail -4 gpurun_out/gpu_tests.log; tail -2 gpurun_out/gpu_tests_fullsize_bits.log; for f in stage_timing stage_timing_tma stage_timing_det; do tail -1 gpurun_out/$f.log; done
for b in a b c; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$b.json"))
    print("$b", {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax")}, d["e2e"]["value"], d["e2e"]["resident_table"]["value"])
except Exception as e: print("$b", e); print(open("gpurun_out/bench_$b.err").read()[-1500:])
PY
done
cat gpurun_out/session.log

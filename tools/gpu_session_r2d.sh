set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing.log 2>&1
SGN_PROJECT_STAGED=1 T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing_staged.log 2>&1
SGN_PROJECT_STAGED=1 T 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "project or binning" > gpurun_out/gpu_tests_staged.log 2>&1
for b in a b c; do
T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_$b.json 2> gpurun_out/bench_$b.err
done
T 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:project_ -o gpurun_out/project_direct python tools/ncu_step.py > /dev/null 2>&1
SGN_PROJECT_STAGED=1 T 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:project_fwd -o gpurun_out/project_staged python tools/ncu_step.py > /dev/null 2>&1
tail -3 gpurun_out/gpu_tests.log; tail -2 gpurun_out/gpu_tests_staged.log; tail -1 gpurun_out/stage_timing.log; tail -1 gpurun_out/stage_timing_staged.log
for b in a b c; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$b.json"))
    print("$b", {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax")}, d["e2e"]["value"], d["e2e"]["resident_table"]["value"])
except Exception as e: print("$b", e); print(open("gpurun_out/bench_$b.err").read()[-1500:])
PY
done
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/session.log

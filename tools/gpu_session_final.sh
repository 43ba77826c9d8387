set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1
T 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
T 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_k20.json 2> gpurun_out/bench_final_k20.err
T 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/ncu_step.py > /dev/null 2>&1
T 500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/step_full python tools/ncu_step.py > /dev/null 2>&1
T 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged_edge or extra_channels or async or deterministic" > gpurun_out/sanitizer_memcheck.log 2>&1
T 400 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged_edge and (forward or backward or project)" > gpurun_out/sanitizer_initcheck.log 2>&1
tail -5 gpurun_out/gpu_tests.log; tail -3 gpurun_out/sanitizer_memcheck.log; tail -3 gpurun_out/sanitizer_initcheck.log
python - <<PY
import json
for f in ("bench_final","bench_final_k20"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax","gpu_launches")}, "e2e", round(d["e2e"]["value"],1), d["clocks"])
    except Exception as e: print(f, e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
ls -la gpurun_out/step_full.ncu-rep; cat gpurun_out/session.log

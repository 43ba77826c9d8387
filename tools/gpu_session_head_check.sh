set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1
T 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
T 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/gpu_tests.log; tail -2 gpurun_out/smoke.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax","gpu_launches")}, "e2e", round(d["e2e"]["value"],1))
PY
cat gpurun_out/session.log

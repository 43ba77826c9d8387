set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 600 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
P=street-gaussians-ns_b200
for v in "" _r96 _r80 _r64; do
  SGN_RASTER_LIB=$PWD/$P/libsgn_raster$v.so T 100 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing$v.log 2>&1
done
T 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
SGN_BENCH_CLOCK_INTERVAL=0 T 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
SGN_BIN_LOCAL=1 T 180 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_bin_local.csv \
    python tools/ncu_step.py > /dev/null 2>&1
tail -5 gpurun_out/gpu_tests.log; for v in "" _r96 _r80 _r64; do tail -1 gpurun_out/stage_timing$v.log; done
for b in a b c; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$b.json"))
    print("$b", {k:d.get(k) for k in ("value","ms_per_step","ms_per_step_median","ms_per_step_max","ms_per_step_argmax")}, d["e2e"]["value"], d["e2e"]["resident_table"]["value"])
    print("   cfg4", json.dumps(d.get("training_step_cfg4"))[:700])
except Exception as e: print("$b", e); print(open("gpurun_out/bench_$b.err").read()[-1500:])
PY
done
cat gpurun_out/session.log

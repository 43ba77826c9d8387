set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
for i in 1 2 3 4 5 6; do
SGN_BENCH_CLOCK_INTERVAL=0 T 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_nosampler_$i.json 2> gpurun_out/bench_nosampler_$i.err
T 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_sampler_$i.json 2> gpurun_out/bench_sampler_$i.err
done
python - <<PY
import json,glob
for kind in ("nosampler","sampler"):
    for f in sorted(glob.glob(f"gpurun_out/bench_{kind}_*.json")):
        try:
            d=json.loads(open(f).read().strip().splitlines()[-1])
            print(kind, round(d["value"],1), round(d["ms_per_step_median"],3), round(d["ms_per_step_max"],2), d["ms_per_step_argmax"], "e2e", round(d["e2e"]["value"],1))
        except Exception as e: print(f, e)
PY
cat gpurun_out/session.log | tail -3

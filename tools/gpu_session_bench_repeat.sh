set -u
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg45 > gpurun_out/bench_rep_$i.json 2> gpurun_out/bench_rep_$i.err
done
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_rep_*.json"))+["gpurun_out/bench_final.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[-10:], round(d["value"],1), round(d["ms_per_step_median"],3), round(d["ms_per_step_max"],2), d["ms_per_step_argmax"], "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["resident_table"]["value"],1))
    except Exception as e: print(f, e)
PY

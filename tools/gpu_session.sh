#!/usr/bin/env bash
# One GPU call that brings back everything a round needs first (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_session.sh'
# Writes into gpurun_out/: the GPU test log, a bench line, the per-stage timing, the config-4 training loop, the
# launch list of one step and of one refinement under ncu (gpu__time_duration only: cheap), and one full ncu capture of a
# refinement (the refinement kernels have no profile yet).  Every step is bounded by its own timeout.
set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
T 420 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
SGN_TEST_EXPERIMENTAL=1 T 240 python -m pytest tests/test_gpu_zz_experimental.py -q -s > gpurun_out/gpu_tests_experimental.log 2>&1
T 240 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
T 120 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing.log 2>&1
SGN_BIN_LOCAL=1 T 120 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing_bin_local.log 2>&1
SGN_BIN_LOCAL=1 T 240 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_bin_local.json 2> gpurun_out/bench_bin_local.err
T 180 python tools/train_cfg4.py --steps 30 --warmup 5 --refine-every 10 --start-step 600 > gpurun_out/train_cfg4.json 2> gpurun_out/train_cfg4.err
T 180 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv \
    python tools/ncu_step.py > /dev/null 2>&1
T 180 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_refine.csv \
    python tools/ncu_refine.py > /dev/null 2>&1
T 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:refine_ -o gpurun_out/refine \
    python tools/ncu_refine.py > /dev/null 2>&1
tail -3 gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests_experimental.log; tail -1 gpurun_out/stage_timing.log; tail -1 gpurun_out/stage_timing_bin_local.log; tail -c 600 gpurun_out/bench.json; cat gpurun_out/session.log

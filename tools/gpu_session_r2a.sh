set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 420 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
T 400 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q > gpurun_out/gpu_tests_fullsize.log 2>&1
SGN_TEST_EXPERIMENTAL=1 T 200 python -m pytest tests/test_gpu_zz_experimental.py -q -s > gpurun_out/gpu_tests_experimental.log 2>&1
T 240 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
T 120 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing.log 2>&1
SGN_BIN_LOCAL=1 T 120 python tools/stage_timing.py --cfg 3 --iters 20 > gpurun_out/stage_timing_bin_local.log 2>&1
T 180 python tools/train_cfg4.py --steps 30 --warmup 5 --refine-every 10 --start-step 600 > gpurun_out/train_cfg4.json 2> gpurun_out/train_cfg4.err
tail -3 gpurun_out/gpu_tests.log; tail -30 gpurun_out/gpu_tests_fullsize.log; tail -15 gpurun_out/gpu_tests_experimental.log; tail -1 gpurun_out/stage_timing.log; tail -1 gpurun_out/stage_timing_bin_local.log; tail -c 600 gpurun_out/bench.json; tail -5 gpurun_out/train_cfg4.err; tail -c 800 gpurun_out/train_cfg4.json; cat gpurun_out/session.log

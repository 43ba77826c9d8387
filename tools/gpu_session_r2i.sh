set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
P=street-gaussians-ns_b200
for v in "" _b20 _b24 _b32; do
  SGN_RASTER_LIB=$PWD/$P/libsgn_raster$v.so T 100 python tools/stage_timing.py --cfg 3 --iters 30 > gpurun_out/stage_timing$v.log 2>&1
done
T 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
for v in "" _b20 _b24 _b32; do tail -1 gpurun_out/stage_timing$v.log; done; tail -3 gpurun_out/gpu_tests.log; cat gpurun_out/session.log

set -u
mkdir -p gpurun_out
T() { local secs=$1; shift; timeout "$secs" "$@"; echo "[exit $?] $*" >> gpurun_out/session.log; }
rm -f gpurun_out/session.log
T 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_parity.py > gpurun_out/gpu_tests.log 2>&1
T 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:blend_ -o gpurun_out/blend_default python tools/ncu_step.py > /dev/null 2>&1
SGN_RASTER_LIB=$PWD/street-gaussians-ns_b200/libsgn_raster_r64.so T 300 ncu --set full --clock-control none --profile-from-start off -k regex:blend_ -o gpurun_out/blend_r64 python tools/ncu_step.py > /dev/null 2>&1
tail -4 gpurun_out/gpu_tests.log; ls -la gpurun_out/*.ncu-rep; cat gpurun_out/session.log

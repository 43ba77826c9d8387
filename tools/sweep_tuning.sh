# A/B of the SGN_TUNE_* execution variants (include/sgn_raster.h) on cfg3: prints blend_fwd / blend_bwd stage times
for t in ${SGN_SWEEP:-0 1 4 5 9 13 12 8}; do SGN_TUNING=$t python tools/stage_timing.py --cfg 3 --iters 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tuning', $t, 'fwd', d['blend_fwd'], 'bwd', d['blend_bwd'], 'sum', d['gpu_sum_ms'])"; done

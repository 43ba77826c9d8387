# A/B of the heavy-first work lists (SGN_HEAVY_FIRST) on cfg3
for ord in 0 1; do echo "heavy_first=$ord"; SGN_HEAVY_FIRST=$ord SGN_SWEEP="${SGN_SWEEP:-12}" bash tools/sweep_tuning.sh; done

#!/bin/bash
# TMA supply-rate experiments (timing build): MMAs skipped, ring depth / CTA count varied.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp2.log; : > $OUT
HERE=$(pwd)
export BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so BHSR_DEBUG_TIMING=1
run() { echo "== $*" >> $OUT; ( env "$@" timeout 120 python tools/probe_conv_tc.py $CASE 0 2>/dev/null | grep -E '"ms"|cycles' ) >> $OUT; }
for CASE in time_exact32_mb2 time_exact32 time_exact32_c160_mb2 time_exact64_c192_mb2 time_fast32 time_fast32_c160_mb2; do
  run CASE=$CASE BHSR_DEBUG_NOMMA=1
done
for CASE in time_exact32_mb2 time_exact32; do
  run CASE=$CASE BHSR_DEBUG_NOMMA=1 BHSR_DEBUG_FORCE_STREAM=1 BHSR_ASTAGES=3
  run CASE=$CASE BHSR_DEBUG_NOMMA=0 BHSR_DEBUG_FORCE_STREAM=1 BHSR_ASTAGES=3
  run CASE=$CASE BHSR_DEBUG_NOMMA=1 BHSR_DEBUG_FORCE_STREAM=1 BHSR_ASTAGES=2
  run CASE=$CASE BHSR_DEBUG_NOMMA=1 PROBE_MAX_CTAS=37
  run CASE=$CASE BHSR_DEBUG_NOMMA=0 PROBE_MAX_CTAS=37
done
CASE=time_exact32 run CASE=time_exact32 BHSR_DEBUG_NOMMA=1 BHSR_DEBUG_FORCE_STREAM=1 BHSR_ASTAGES=4
cat $OUT | cut -c1-600

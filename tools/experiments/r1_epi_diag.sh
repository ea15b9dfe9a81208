#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/exp_epi.log; : > $OUT
HERE=$(pwd)
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so
for m in 4 3; do
  echo "== time_exact32_mb2 NOMMA=$m" >> $OUT
  BHSR_DEBUG_NOMMA=$m timeout 40 python tools/probe_conv_tc.py time_exact32_mb2 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-420 >> $OUT
done
cat $OUT

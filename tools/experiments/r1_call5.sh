#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp5.log; : > $OUT
HERE=$(pwd)
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so
for c in time_exact32_mb2_nb16 time_exact32_mb2_nb32 time_exact32_c160_mb2_nb16 time_exact32_mb2_ct64 time_exact64_c192_mb2_nb16; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
  echo "== $c NOMMA" >> $OUT
  BHSR_DEBUG_NOMMA=1 timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
cat $OUT | cut -c1-560

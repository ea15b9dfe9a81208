#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp_slot.log; : > $OUT
HERE=$(pwd)
for c in time_exact32_mb2 time_exact32_c160_mb2 time_fast32; do
  echo "== $c" >> $OUT
  BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles|frac_bad' | cut -c1-330 >> $OUT
done
cat $OUT
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2

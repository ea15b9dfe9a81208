#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp.log; : > $OUT
for c in time_exact32 time_exact32_mb2 time_exact64_c192_mb2 time_fast32; do
  for m in 0 1 2 3; do
    echo "== $c mode $m" >> $OUT
    BHSR_DEBUG_TIMING=1 timeout 120 python tools/probe_conv_tc.py $c $m 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
  done
done
cat $OUT | cut -c1-300

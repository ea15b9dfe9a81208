#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp.log; : > $OUT
for c in time_fast32 time_fast32_mb1 time_exact32 time_fast64_c192 time_exact64_c192; do
    echo "== $c" >> $OUT
    BHSR_DEBUG_TIMING=1 timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
cat $OUT

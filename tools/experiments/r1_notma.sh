#!/bin/bash
# MMA stream of the dx kernel with and without TMA traffic (timing build, BHSR_DEBUG_NOMMA=2 keeps
# stale activation tiles after the first fill): separates shared-memory-port contention from issue cost.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp_notma.log; : > $OUT
HERE=$(pwd)
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so
for c in time_exact32_mb2 time_exact32_c160_mb2 time_fast32; do
  for m in 0 2; do
    echo "== $c NOMMA=$m" >> $OUT
    BHSR_DEBUG_NOMMA=$m timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
  done
done
cat $OUT

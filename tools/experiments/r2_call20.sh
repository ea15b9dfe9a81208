#!/bin/bash
# Round 2, call 20: issue blocks without barrier probes (NOMMA=3) and the cycles of the post-issue section
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c20_probes.log; : > $OUT
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so
for c in time_fast32 time_fast32_c160_mb2; do
  for m in 0 3; do
    echo "== $c DXS_MB=2 timing NOMMA=$m" >> $OUT
    BHSR_DXS_MB=2 BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
  done
done
for c in time_fast32_c160_mb4 time_fast32_c64_mb4; do
  for m in 0 3; do
    echo "== $c (dxs) timing NOMMA=$m" >> $OUT
    BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
  done
done
cat $OUT

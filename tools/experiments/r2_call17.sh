#!/bin/bash
# Round 2, call 17: what bounds the exact dx layers with many chunks (conv4): TMA supply vs MMA stream
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c17_exact_c160.log; : > $OUT
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so
for c in time_exact32_c160_mb2 time_exact32_c96_mb2; do
  for m in 0 1 2 4; do
    echo "== $c timing NOMMA=$m" >> $OUT
    BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
  done
done
for ns in 2 3; do
  echo "== time_exact32_c160_mb2 ASTAGES=$ns" >> $OUT
  BHSR_ASTAGES=$ns timeout 60 python tools/probe_conv_tc.py time_exact32_c160_mb2 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
done
echo "== time_exact64_c192_mb2 (pair) timing" >> $OUT
timeout 60 python tools/probe_conv_tc.py time_exact64_c192_mb2 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
cat $OUT

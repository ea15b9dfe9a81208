#!/bin/bash
# Round 2, call 15: where the dx kernel stands after the compact epilogue: CTA pairs again, fast-mode diagnostics
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c15_layers.log; : > $OUT
for c in time_exact32_mb2 time_exact32_c96_mb2 time_exact32_c128_mb2 time_exact32_c160_mb2; do
  for pr in 0 1; do
    echo "== $c DX_PAIR=$pr" >> $OUT
    BHSR_DX_PAIR=$pr timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
  done
done
for c in time_fast32 time_fast32_c160_mb2 time_fast64_c192 time_exact64_c192_mb2; do
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
done
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so
for c in time_fast32 time_fast32_c160_mb2 time_exact32_mb2; do
  for m in 0 2 4; do
    echo "== $c timing NOMMA=$m" >> $OUT
    BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
  done
done
for m in 0 4; do
  echo "== time_exact32_mb2 DX_PAIR=1 timing NOMMA=$m" >> $OUT
  BHSR_DX_PAIR=1 BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py time_exact32_mb2 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
done
cat $OUT

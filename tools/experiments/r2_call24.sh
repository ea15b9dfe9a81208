#!/bin/bash
# Round 2, call 24: where the dx epilogue's cycles go (drain / exchange barrier / shuffle combine / finish+stores)
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c24_epistages.log; : > $OUT
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so BHSR_DEBUG_NOMMA=6
for c in time_fast32 time_fast32_c160_mb2; do
  echo "== $c DXS_MB=2 epilogue stages" >> $OUT
  BHSR_DXS_MB=2 timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
  echo "== $c DXS_MB=2 LEAN epilogue stages" >> $OUT
  BHSR_DXS_LEAN=1 BHSR_DXS_MB=2 timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
done
cat $OUT

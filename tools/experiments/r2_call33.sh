#!/bin/bash
# Round 2, call 33: MMA-warp counters of the lean exact issuer (total / waits per tile): is the exact MMA stream execution-bound?
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c33_lean_exact_counters.log; : > $OUT
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so
for c in time_exact32_mb2 time_exact32_c160_mb2; do
  for lean in 0 1; do
    for m in 0 9; do
      echo "== $c DX_LEAN=$lean NOMMA=$m" >> $OUT
      BHSR_DX_LEAN=$lean BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
    done
  done
done
cat $OUT

#!/bin/bash
# Round 2, call 16: single-accumulator tall-tile dx kernel (conv_dxs), fast numerics: correctness + layer times
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c16_dxs.log; : > $OUT
for mb in 3 4; do
  for c in fast32_mb$mb fast32_c160_mb$mb fast32_c96_w130_mb$mb fast32_c32_h7_mb$mb; do
    echo "== $c" >> $OUT
    timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror|timeout' | cut -c1-300 >> $OUT
  done
done
echo "== fast32_mb2 via BHSR_DXS_MB=2" >> $OUT
BHSR_DXS_MB=2 timeout 60 python tools/probe_conv_tc.py fast32_mb2 0 2>&1 | grep -E 'max_abs_err|rror|timeout' | cut -c1-300 >> $OUT
for c in time_fast32 time_fast32_c96_mb2 time_fast32_c128_mb2 time_fast32_c160_mb2; do
  echo "== $c (old kernel)" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
  echo "== $c BHSR_DXS_MB=2" >> $OUT
  BHSR_DXS_MB=2 timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
done
for mb in 3 4; do
  for cc in 64 96 128 160; do
    echo "== time_fast32_c${cc}_mb$mb" >> $OUT
    timeout 60 python tools/probe_conv_tc.py time_fast32_c${cc}_mb$mb 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
  done
done
cat $OUT

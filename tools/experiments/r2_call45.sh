#!/bin/bash
# Round 2, call 45: does the trunk pay for scattering 64-byte results into 384-byte concat records?  compact output plane vs concat slice
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c45_out_layout.log; : > $OUT
for c in time_exact32_mb2 time_exact32_c64_o32 time_exact32_c96_mb2 time_exact32_c96_o32 time_exact32_c128_mb2 time_exact32_c128_o32 time_exact32_c160_mb2 time_exact32_c160_o32 time_fast32 time_fast32_c64_o32; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"|rror' | cut -c1-200 >> $OUT
done
cat $OUT

#!/bin/bash
# Round 2, call 40: lean MMA issuer on the head's one-chunk layers (K = 144): 16 -> 16 and 32 -> 16 padded to 32 outputs at 256x256, B = 32
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c40_lean_head.log; : > $OUT
for lean in 0 1 2; do
  echo "== time_exact32_c16_256 DX_LEAN=$lean" >> $OUT
  BHSR_DX_LEAN=$lean timeout 120 python tools/probe_conv_tc.py time_exact32_c16_256 0 2>&1 | grep -E '"ms"|max_abs_err|rror' | cut -c1-200 >> $OUT
done
cat $OUT

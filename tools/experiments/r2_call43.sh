#!/bin/bash
# Round 2, call 43: L2 promotion of the activation maps for the head-sized layer (contiguous 64-byte pixel records, HBM-sourced)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c43_l2promo_head.log; : > $OUT
for pr in 0 1 2 3; do
  echo "== time_exact16_c16_256 L2PROMO=$pr" >> $OUT
  BHSR_L2PROMO=$pr timeout 120 python tools/probe_conv_tc.py time_exact16_c16_256 0 2>&1 | grep -E '"ms"|rror' | cut -c1-200 >> $OUT
done
cat $OUT

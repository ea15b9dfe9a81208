#!/bin/bash
# Round 2, call 44: 16-channel chunks (SWIZZLE_32B) for the 16 -> 16 layers on 16-channel planes: bring-up + time
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c44_c16.log; : > $OUT
for c in exact16_t16 exact16_t16_small exact16_t48_off16 time_exact16_t16_256 time_exact16_c16_256_o32 time_exact16_c16_256; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"|max_abs_err|rror|trap|timeout' | cut -c1-260 >> $OUT
done
cat $OUT

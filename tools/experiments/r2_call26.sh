#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/r2c26_pair_small_cin.log; : > $OUT
for cc in 16 32 64 96; do for nb in 2 3; do for sfx in "" _nchw; do
  c=exact64_c${cc}_nb${nb}${sfx}
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror' | cut -c1-300 >> $OUT
done; done; done
cat $OUT

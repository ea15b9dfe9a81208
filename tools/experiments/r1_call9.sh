#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp9.log; : > $OUT
HERE=$(pwd)
for promo in 3 2 1 0; do
  for c in time_exact32_mb2 time_exact32_c160_mb2 time_exact64_c192_mb2; do
    echo "== $c promo=$promo" >> $OUT
    BHSR_L2PROMO=$promo BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
  done
  echo "== bench promo=$promo" >> $OUT
  BHSR_L2PROMO=$promo timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" >> $OUT
done
cat $OUT | cut -c1-420

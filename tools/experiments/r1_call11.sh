#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp11.log; : > $OUT
for c in time_exact32_c96 time_exact32_c96_mb2 time_exact32_c128 time_exact32_c128_mb2 time_exact32_c160 time_exact32_c160_mb2 time_fast32 time_fast32_mb1; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' >> $OUT
done
cat $OUT | cut -c1-300
for c in 3 5; do
  timeout 900 python tools/bench_configs.py --config $c > gpurun_out/config$c.log 2>&1; echo "config $c rc=$?"; tail -1 gpurun_out/config$c.log | cut -c1-500
done
bash tools/run_head_profile.sh

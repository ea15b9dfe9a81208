#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp8.log; : > $OUT
HERE=$(pwd)
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so
for c in time_fast32 time_fast32_ct64 time_exact32_c32_ct32 time_exact32_c32_ct192 time_fast64_ct64 time_fast64_ct192; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
  echo "== $c NOMMA" >> $OUT
  BHSR_DEBUG_NOMMA=1 timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
cat $OUT | cut -c1-560
unset BHSR_DEBUG_TIMING BHSR_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log; tail -4 gpurun_out/pytest_gpu8.log

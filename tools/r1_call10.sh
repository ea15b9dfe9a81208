#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp10.log; : > $OUT
HERE=$(pwd)
for c in time_exact32_mb2 time_exact32 time_exact32_c160_mb2 time_exact64_c192_mb2 time_fast32_c160_mb2 time_fast64_c192; do
  echo "== $c" >> $OUT
  BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
cat $OUT | cut -c1-420
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10.log; tail -4 gpurun_out/pytest_gpu10.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench10.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench10.log
tail -2 gpurun_out/bench10.log | cut -c1-1800

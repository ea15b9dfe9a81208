#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp16.log; : > $OUT
HERE=$(pwd)
for c in time_exact32_mb2 time_exact64_c192_mb2; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
cat $OUT | cut -c1-420
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu16.log; tail -4 gpurun_out/pytest_gpu16.log
timeout 600 python bench.py --steps 10 > gpurun_out/bench16.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench16.log
tail -2 gpurun_out/bench16.log | cut -c1-1800

#!/bin/bash
# GPU-side check run: warm the page cache, run the gpu tests, the conv timing probes and a short bench.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.get_device_name(0))" > gpurun_out/warm.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
if [ -n "$PROBE_CASES" ]; then
  OUT=gpurun_out/exp.log; : > $OUT
  for c in $PROBE_CASES; do
    echo "== $c" >> $OUT
    BHSR_DEBUG_TIMING=1 timeout 180 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
  done
  cat $OUT | cut -c1-330
fi
if [ -n "$BENCH" ]; then
  timeout 900 python bench.py $BENCH > gpurun_out/bench.log 2>&1
  echo "bench rc=$?" >> gpurun_out/bench.log
  tail -5 gpurun_out/bench.log | cut -c1-3000
fi

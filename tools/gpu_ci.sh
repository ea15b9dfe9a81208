#!/bin/bash
# GPU-side check run: warm the page cache, run the gpu tests, the conv timing probes and a short bench.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.get_device_name(0))" > gpurun_out/warm.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
if [ -n "$PROBE_CASES" ]; then
  CASES="$PROBE_CASES" MODES=0 bash tools/run_probe.sh > gpurun_out/probe_cases.log 2>&1
  grep -E '"ms"|rc=' gpurun_out/probe_cases.log
fi
if [ -n "$BENCH" ]; then
  timeout 900 python bench.py $BENCH > gpurun_out/bench.log 2>&1
  echo "bench rc=$?" >> gpurun_out/bench.log
  tail -5 gpurun_out/bench.log | cut -c1-3000
fi

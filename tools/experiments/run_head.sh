#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_head_gpu.py tests/test_rrdbnet_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for c in 3 5; do
  timeout 900 python tools/bench_configs.py --config $c > gpurun_out/config$c.log 2>&1; echo "config $c rc=$?"; tail -1 gpurun_out/config$c.log | cut -c1-400
done

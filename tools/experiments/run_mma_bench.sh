#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/mma_bench > gpurun_out/mma_bench.log 2>&1
echo "rc=$?" >> gpurun_out/mma_bench.log
cat gpurun_out/mma_bench.log
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -x -q > gpurun_out/pytest_head.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_head.log
tail -40 gpurun_out/pytest_head.log

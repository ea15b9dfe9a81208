#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_head_gpu.py -m gpu -q -x -k "wgrad_tc_kernel_vs_fp64 and 32-16-1" > gpurun_out/r2c8_k1.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_k1.log
grep -E "passed|failed|FAILED|outside|rc=|Error" gpurun_out/r2c8_k1.log | head -20

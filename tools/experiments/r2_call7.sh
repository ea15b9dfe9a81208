#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_head_gpu.py -m gpu -q -k "conv_tc_train" > gpurun_out/r2c7_convtrain.log 2>&1; echo "rc=$?" >> gpurun_out/r2c7_convtrain.log
grep -E "passed|failed|FAILED|outside|rc=|Error" gpurun_out/r2c7_convtrain.log | head -20
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_head_gpu.py -m gpu -q -x -k "wgrad_tc_kernel_vs_fp64 and 16-16-3" > gpurun_out/r2c7_sanitizer.log 2>&1
grep -E "=========" gpurun_out/r2c7_sanitizer.log | head -40

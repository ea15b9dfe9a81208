#!/bin/bash
# Round 2, call 35: training step with the smp part forked onto a second stream inside the CUDA graph
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -q -k "overlap or channels_last" > gpurun_out/r2c35_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c35_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c35_pytest.log | head

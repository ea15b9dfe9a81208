#!/bin/bash
# Round 2, call 28: row N3 tests (all), SR fine-tune step time (config 6)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -s -k "23block_backward" > gpurun_out/r2c28_n3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c28_n3.log
grep -E "passed|failed|FAILED|rel-L2|rc=|Error|error|assert" gpurun_out/r2c28_n3.log | head -40

#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 300 python tools/debug_n3c.py > gpurun_out/r2c27_debug_n3.log 2>&1
cat gpurun_out/r2c27_debug_n3.log | tail -80

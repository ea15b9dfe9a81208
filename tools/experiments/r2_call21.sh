#!/bin/bash
# Round 2, call 21: all gpu tests with the vectorised BatchNorm element-wise kernels and the dxs kernel; config 3 time
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2c21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c21_pytest.log
grep -E "passed|failed|FAILED|outside|rc=|Error|s call" gpurun_out/r2c21_pytest.log | head -30
timeout 900 python tools/bench_configs.py --config 3 --steps 8 --warmup 3 > gpurun_out/r2c21_cfg3.log 2>&1; tail -1 gpurun_out/r2c21_cfg3.log | cut -c1-500

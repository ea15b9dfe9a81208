#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2c14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c14_pytest.log
grep -E "passed|failed|FAILED|outside|rc=|Error|s call" gpurun_out/r2c14_pytest.log | head -30
timeout 900 python tools/bench_configs.py --config 5 --grids 2560 > gpurun_out/r2c14_cfg5.log 2>&1; tail -1 gpurun_out/r2c14_cfg5.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

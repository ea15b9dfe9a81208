#!/bin/bash
# Round 2, call 46: wgrad_tc_kernel<32> (64-channel input groups for the 32-output RDB convs): N3 + head tests, SR fine-tune step
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q -k "backward or wgrad or finetune or optimize_parameters or generator_step or head" > gpurun_out/r2c46_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c46_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c46_pytest.log | head
timeout 900 python tools/bench_configs.py --config 6 --steps 3 --warmup 1 > gpurun_out/r2c46_cfg6.log 2>&1; tail -1 gpurun_out/r2c46_cfg6.log | cut -c1-300

#!/bin/bash
# Round 2, call 28: row N3 tests (all), SR fine-tune step time (config 6)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -s -k "backward_vs_oracle or finetune or error_behaviour or optimize_parameters or generator_step" > gpurun_out/r2c28_n3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c28_n3.log
grep -E "passed|failed|FAILED|rel-L2|rc=|Error|error|assert" gpurun_out/r2c28_n3.log | head -40
timeout 900 python tools/bench_configs.py --config 6 --steps 3 --warmup 1 > gpurun_out/r2c28_cfg6.log 2>&1; tail -2 gpurun_out/r2c28_cfg6.log | cut -c1-600
timeout 900 python tools/bench_configs.py --config 6 --steps 3 --warmup 1 --batch 4 > gpurun_out/r2c28_cfg6_b4.log 2>&1; tail -1 gpurun_out/r2c28_cfg6_b4.log | cut -c1-600
BHSR_SR_GRAPH=0 timeout 900 python tools/bench_configs.py --config 6 --steps 3 --warmup 1 > gpurun_out/r2c28_cfg6_eager.log 2>&1; tail -1 gpurun_out/r2c28_cfg6_eager.log | cut -c1-400

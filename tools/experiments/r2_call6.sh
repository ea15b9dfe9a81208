#!/bin/bash
# Round 2, call 6: tensor-core training head bring-up (to_planes, conv_tc NCHW/accumulate, wgrad_tc)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -q -k "wgrad_tc or conv_tc_train" > gpurun_out/r2c6_unit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c6_unit.log
grep -E "passed|failed|Error|error|outside|rc=" gpurun_out/r2c6_unit.log | head -30
timeout 1200 python -m pytest tests/test_head_gpu.py -m gpu -q > gpurun_out/r2c6_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c6_head.log
grep -E "passed|failed|FAILED|outside|rc=" gpurun_out/r2c6_head.log | head -30
timeout 900 python tools/bench_configs.py --config 3 --steps 5 --warmup 2 > gpurun_out/r2c6_cfg3.log 2>&1; tail -1 gpurun_out/r2c6_cfg3.log | cut -c1-400
BHSR_HEAD_TC_TRAIN=0 timeout 900 python tools/bench_configs.py --config 3 --steps 5 --warmup 2 > gpurun_out/r2c6_cfg3_cudacore.log 2>&1; tail -1 gpurun_out/r2c6_cfg3_cudacore.log | cut -c1-400

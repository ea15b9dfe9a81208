#!/bin/bash
# Round 2, call 22: launch list of two config-3 training steps after the element-wise kernel rewrite
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 9000 --csv --log-file gpurun_out/r2c22_launches_cfg3.csv python tools/bench_configs.py --config 3 --steps 2 --warmup 2 > gpurun_out/r2c22_ncu.log 2>&1
tail -2 gpurun_out/r2c22_ncu.log | cut -c1-300

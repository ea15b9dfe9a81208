#!/bin/bash
# Round 2, call 47: launch list of one SR fine-tune generator step (2-block generator, B = 12, eager)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
BHSR_SR_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1400 --csv --log-file gpurun_out/r2c47_launches_n3.csv python tools/bench_configs.py --config 6 --steps 1 --warmup 1 --num-block 2 > gpurun_out/r2c47_ncu.log 2>&1
tail -1 gpurun_out/r2c47_ncu.log | cut -c1-200

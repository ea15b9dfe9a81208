#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
# inference: one sweep = 1 batch of 128 (after a 1-batch warm-up sweep); skip RRDB trunk by using 1 block
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg5.csv \
   python tools/bench_configs.py --config 5 --grids 128 --batch 128 --num-block 1 > gpurun_out/launches_cfg5.log 2>&1
echo "cfg5 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv \
   python tools/bench_configs.py --config 3 --steps 1 --warmup 1 --num-block 1 > gpurun_out/launches_cfg3.log 2>&1
echo "cfg3 rc=$?"
ls -la gpurun_out/launches_cfg*.csv

#!/bin/bash
# Launch list of ONE training step (config 3, B = 32, 23-block trunk, eager launches) with the final library
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3_final.csv \
   python tools/bench_configs.py --config 3 --steps 1 --warmup 1 > gpurun_out/launches_cfg3_final.log 2>&1
echo "cfg3 rc=$?"
tail -2 gpurun_out/launches_cfg3_final.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/launches_cfg3_final.csv 40 | tee gpurun_out/launches_cfg3_final.md | head -50

#!/bin/bash
# Round 2, call 41: ncu --set full of the head-sized dx layer (16 -> 16 at 256x256, B = 32), 16-output and padded-to-32 kernels
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for c in time_exact16_c16_256 time_exact32_c16_256; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_dx_kernel -s 2 -c 1 -f -o gpurun_out/prof_$c python tools/probe_conv_tc.py $c 0 > gpurun_out/r2c41_$c.log 2>&1
  echo "$c rc=$?"
done
ls -la gpurun_out/prof_time_exact*.ncu-rep

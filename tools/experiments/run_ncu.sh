#!/bin/bash
# ncu captures of the tensor-core conv kernel on representative layer shapes.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for c in ${CASES:-time_fast32 time_exact32 time_exact64_c192}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 6 -c 2 \
     -f -o gpurun_out/prof_$c python tools/probe_conv_tc.py $c 0 > gpurun_out/ncu_$c.log 2>&1
  echo "$c rc=$?"
done
ls -la gpurun_out/*.ncu-rep

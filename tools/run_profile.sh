#!/bin/bash
# Round profile: per-launch device times of one bench step + one full capture of the RDB conv kernels.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
NUM=${NUMERICS:-exact}
RX='regex:conv_tc_kernel|conv_dx_kernel|conv_pair_kernel|conv3x3_'
# launch list: 3 warm-up forwards x 356 conv launches are skipped, one whole step is listed
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" \
   -s 1068 -c 356 --csv --log-file gpurun_out/launches_$NUM.csv \
   python bench.py --steps 1 --warmup 3 --numerics $NUM --no-cpu-baseline --no-secondary --no-train --no-graph > gpurun_out/launches_$NUM.log 2>&1
echo "launch list rc=$?"
# full capture of 10 consecutive trunk convs (two RDBs' worth: conv1..conv5 appear in order)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel|conv_dx_kernel|conv_pair_kernel" -s 1100 -c 10 \
   -f -o gpurun_out/prof_rdb_$NUM python bench.py --steps 1 --warmup 3 --numerics $NUM --no-cpu-baseline --no-secondary --no-train --no-graph > gpurun_out/prof_rdb_$NUM.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv

#!/bin/bash
# Round profile: per-launch device times of one bench step + one full capture of the top kernels.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
NUM=${NUMERICS:-exact}
# launch list: skip pack kernels + 3 warm-up forwards (357 launches each), capture one step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel|conv3x3_' \
   -s 1071 -c 357 --csv --log-file gpurun_out/launches_$NUM.csv \
   python bench.py --steps 1 --warmup 3 --numerics $NUM --no-cpu-baseline --no-secondary > gpurun_out/launches_$NUM.log 2>&1
echo "launch list rc=$?"
# full capture: rdb conv1 (N=32,K=576), conv5 (N=64,K=1728), conv_hr (256x256) of one forward
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1071 -c 5 \
   -f -o gpurun_out/prof_rdb_$NUM python bench.py --steps 1 --warmup 3 --numerics $NUM --no-cpu-baseline --no-secondary > gpurun_out/prof_rdb_$NUM.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv

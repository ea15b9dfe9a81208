#!/bin/bash
# Round 2, call 30: ncu --set full of the head's training kernels (one config-3 step, 1-block trunk, B=32)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
RX='regex:wgrad_tc_kernel|head_to_planes_kernel|head_from_planes_kernel|head_split_nchw_kernel|bn_bwd_apply_v4|bn_bwd_reduce_v4|affine_add_relu_v4|wgrad_reduce_kernel|weighted_mse'
timeout 1200 ncu --set full --clock-control none -k "$RX" -s 200 -c 40 -f -o gpurun_out/prof_head_train \
   python tools/bench_configs.py --config 3 --steps 1 --warmup 1 --num-block 1 > gpurun_out/r2c30_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_head_train.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active > gpurun_out/r2c30_head_kernels.csv 2>/dev/null
head -3 gpurun_out/r2c30_head_kernels.csv | cut -c1-400
ls -la gpurun_out/prof_head_train.ncu-rep

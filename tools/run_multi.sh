#!/bin/bash
# 2-GPU validation of the launch contract: bench.py under torchrun (weak scaling, no collective on
# the feature path) and the DP training step (config 4: one NCCL all-reduce per step).
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.device_count())" > gpurun_out/multi.log 2>&1
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 >> gpurun_out/multi.log 2>&1
echo "bench rc=$?" >> gpurun_out/multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 >> gpurun_out/multi.log 2>&1
echo "reference rc=$?" >> gpurun_out/multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   tools/bench_configs.py --config 4 --steps 3 --warmup 2 >> gpurun_out/multi.log 2>&1
echo "config4 rc=$?" >> gpurun_out/multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
   tools/bench_configs.py --config 5 --grids 2560 >> gpurun_out/multi.log 2>&1
echo "config5 rc=$?" >> gpurun_out/multi.log
grep -E "^\{|rc=" gpurun_out/multi.log | cut -c1-500

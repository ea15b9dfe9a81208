#!/bin/bash
# 2-GPU contract check with the final code: both bench arms under torchrun (forward weak scaling + DP training step with
# its NCCL all-reduce in the `train` sub-record)
mkdir -p gpurun_out
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi2_final.log 2>&1
echo "bench rc=$?" >> gpurun_out/multi2_final.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 >> gpurun_out/multi2_final.log 2>&1
echo "reference rc=$?" >> gpurun_out/multi2_final.log
grep -E "^\{|rc=" gpurun_out/multi2_final.log | cut -c1-700
python - <<'PY'
import json
for l in open('gpurun_out/multi2_final.log'):
    if l.startswith('{'):
        d=json.loads(l)
        if d.get('impl')=='reference': print('reference', d['value'], d['cpu_baseline']['kind'])
        else: print('value',d['value'],'e2e',d['e2e']['value'],'train',d['train']['value'],d['train']['ms_per_step'],d['train']['collective'])
PY

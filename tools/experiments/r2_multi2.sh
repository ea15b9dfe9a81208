#!/bin/bash
# 2-GPU contract check: both arms of bench.py under torchrun exactly as the driver launches them
mkdir -p gpurun_out
nvidia-smi -L
python -c "import torch; print(torch.cuda.device_count())"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2m2_ref.log 2>&1; echo "ref rc=$?"
grep '^{' gpurun_out/r2m2_ref.log | tail -1 | cut -c1-300
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m2_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
ls=[l for l in open('gpurun_out/r2m2_bench.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d.get('train',{})
    print('n_gpus',d['n_gpus'],'fwd',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
    print('train',t.get('value'),t.get('ms_per_step'),'eager',t.get('eager_ms_per_step'),t.get('launch'),t.get('collective'),'loss',t.get('loss'))
else:
    print(open('gpurun_out/r2m2_bench.log').read()[-3000:])
PY

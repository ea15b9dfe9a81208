#!/bin/bash
# Round 2, call 37: fused Adam / cuDNN benchmark mode for the stock-PyTorch part of the training step
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  BHSR_FUSED_ADAM=$1 BHSR_CUDNN_BENCHMARK=$2 timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c37_bench_$1$2.log 2>&1
  python - <<PY
import json
ls=[l for l in open('gpurun_out/r2c37_bench_$1$2.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d['train']
    print('fused_adam $1 cudnn_benchmark $2: train',round(t['value'],1),round(t['ms_per_step'],2),'eager',round(t.get('eager_ms_per_step',0),2),'loss',round(t['loss'],3),t['launch'][:12],'clocks',d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/r2c37_bench_$1$2.log').read()[-1500:])
PY
done

#!/bin/bash
# Round 2, call 36: stock-PyTorch smp part in channels_last inside the training step
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for nhwc in 0 1; do
  BHSR_SMP_NHWC=$nhwc timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c36_bench_nhwc$nhwc.log 2>&1
  python - <<PY
import json
ls=[l for l in open('gpurun_out/r2c36_bench_nhwc$nhwc.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d['train']
    print('nhwc $nhwc: fwd',round(d['value'],1),'train',round(t['value'],1),round(t['ms_per_step'],2),'eager',round(t.get('eager_ms_per_step',0),2),'loss',t['loss'],t['launch'][:30],'clocks',d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/r2c36_bench_nhwc$nhwc.log').read()[-1500:])
PY
done

#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for b in 1 0; do
BHSR_CUDNN_BENCHMARK=$b timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c12_bench_cb$b.log 2>&1
python - <<PY
import json
ls=[l for l in open('gpurun_out/r2c12_bench_cb$b.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d.get('train',{})
    print('cudnn.benchmark=$b fwd',d['value'],'train',t.get('value'),t.get('ms_per_step'),'eager',t.get('eager_ms_per_step'),t.get('launch')[:20])
else:
    print(open('gpurun_out/r2c12_bench_cb$b.log').read()[-1500:])
PY
done

#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_head_gpu.py -m gpu -q -k "wgrad_tc_kernel" > gpurun_out/r2c9_wgrad.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_wgrad.log
grep -E "passed|failed|FAILED|outside|rc=|Error" gpurun_out/r2c9_wgrad.log | head -20
timeout 1200 python -m pytest tests/test_head_gpu.py -m gpu -q > gpurun_out/r2c9_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c9_head.log
grep -E "passed|failed|FAILED|outside|rc=" gpurun_out/r2c9_head.log | head -30
timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c9_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2c9_bench.log
python - <<'PY'
import json
ls=[l for l in open('gpurun_out/r2c9_bench.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d.get('train',{})
    print('fwd',d['value'],'train',t.get('value'),t.get('ms_per_step'),'eager',t.get('eager_ms_per_step'),t.get('launch'),'loss',t.get('loss'))
else:
    print(open('gpurun_out/r2c9_bench.log').read()[-2500:])
PY

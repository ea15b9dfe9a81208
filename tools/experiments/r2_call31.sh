#!/bin/bash
# Round 2, call 31: to_planes (float4 shared stores, packed conversions) and coalesced wgrad_reduce: tests + configs 3 / 6
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c31_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c31_pytest.log | head -20
timeout 900 python tools/bench_configs.py --config 3 --steps 8 --warmup 3 > gpurun_out/r2c31_cfg3.log 2>&1; tail -1 gpurun_out/r2c31_cfg3.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c31_bench.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c31_bench.log') if l.startswith('{')][-1]); t=d['train']
print('fwd',d['value'],'train',t['value'],t['ms_per_step'],'eager',t.get('eager_ms_per_step'),'clocks',d['clocks']['sm_mhz'])
PY
timeout 900 python tools/bench_configs.py --config 6 --steps 3 --warmup 1 > gpurun_out/r2c31_cfg6.log 2>&1; tail -1 gpurun_out/r2c31_cfg6.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -k "cuda_graph or error_behaviour or full_batch_properties" > gpurun_out/r2c13_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r2c13_graph.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/r2c13_graph.log | head
for g in "" "--no-graph"; do
timeout 900 python bench.py --no-cpu-baseline --no-train $g > gpurun_out/r2c13_bench$g.log 2>&1
python - <<PY
import json
ls=[l for l in open('gpurun_out/r2c13_bench$g.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1])
    print('graph="$g" value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'fast',d.get('other_numerics',{}).get('value'),'full d2h',d.get('e2e_full_output_d2h',{}).get('value'), 'clk', d['clocks'])
else:
    print(open('gpurun_out/r2c13_bench$g.log').read()[-1500:])
PY
done

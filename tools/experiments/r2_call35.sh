#!/bin/bash
# Round 2, call 35: training step with the smp part forked onto a second stream inside the CUDA graph
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -q -k "overlap or full_pipeline or epoch_harness" > gpurun_out/r2c35_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c35_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c35_pytest.log | head
for ov in 1 2; do
  BHSR_TRAIN_PREFETCH=$((ov-1)) timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c35_bench_ov$ov.log 2>&1
  python - <<PY
import json
ls=[l for l in open('gpurun_out/r2c35_bench_ov$ov.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d['train']
    print('prefetch $((ov-1)): fwd',round(d['value'],1),'train',round(t['value'],1),round(t['ms_per_step'],2),'eager',round(t.get('eager_ms_per_step',0),2),t['launch'][:40],'clocks',d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/r2c35_bench_ov$ov.log').read()[-1500:])
PY
done

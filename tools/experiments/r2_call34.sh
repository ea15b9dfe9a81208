#!/bin/bash
# Round 2, call 34: lean exact issuer + last chunk block-major across both phases (BHSR_DX_LEAN=2)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
BHSR_DX_LEAN=2 timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -k "dx_kernel or rdb_and_rrdb or 2block or full_batch_properties or conv_tc_vs_oracle" > gpurun_out/r2c34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c34_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c34_pytest.log | head -20
OUT=gpurun_out/r2c34_lean2_exact.log; : > $OUT
for c in time_exact32_mb2 time_exact32_c96_mb2 time_exact32_c128_mb2 time_exact32_c160_mb2; do
  for lean in 0 1 2; do
    echo "== $c DX_LEAN=$lean" >> $OUT
    BHSR_DX_LEAN=$lean timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
  done
done
cat $OUT
for lean in 0 2; do
  BHSR_DX_LEAN=$lean timeout 900 python bench.py --no-cpu-baseline --no-secondary --no-train --steps 10 --warmup 3 > gpurun_out/r2c34_bench_lean$lean.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2c34_bench_lean$lean.log') if l.startswith('{')][-1])
print('lean $lean: fwd',round(d['value'],1),'ms',round(d['ms_per_step'],2),'clocks',d['clocks']['sm_mhz'], [round(k['us'],1) for k in d['roofline']['kernels']])
PY
done

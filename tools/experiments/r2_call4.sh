#!/bin/bash
# Round 2, call 4: rolled issue / epilogue loops (code-size engineering): correctness + layer timings + bench
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1500 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -x > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
tail -5 gpurun_out/r2c4_pytest.log
OUT=gpurun_out/r2c4_layers.log; : > $OUT
for c in time_exact32_mb2 time_exact32_c96_mb2 time_exact32_c128_mb2 time_exact32_c160_mb2 time_exact64_c192_mb2 time_fast32 time_fast32_c160_mb2 time_fast64_c192; do
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
done
cat $OUT
timeout 600 python bench.py --no-cpu-baseline --no-train > gpurun_out/r2c4_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2c4_bench.log
tail -2 gpurun_out/r2c4_bench.log | cut -c1-400
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c4_bench.log') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'fast',d.get('other_numerics',{}).get('value'))
for k in d['roofline']['kernels']: print(k['layer'], round(k['us'],1), round(k['tflops'],1))
PY

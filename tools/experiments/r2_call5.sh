#!/bin/bash
# Round 2, call 5: compact (staged, rolled) dx epilogue + rolled per-tap / pair epilogue loops; issue loops unrolled again
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1500 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q > gpurun_out/r2c5_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2c5_pytest.log
tail -8 gpurun_out/r2c5_pytest.log
if [ $rc -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -x -k "test_conv_tc_vs_oracle" > gpurun_out/r2c5_sanitizer.log 2>&1
  grep -E "Invalid|at 0x|by thread|in .*\.cu" gpurun_out/r2c5_sanitizer.log | head -30
fi
OUT=gpurun_out/r2c5_layers.log; : > $OUT
for c in time_exact32_mb2 time_exact32_c96_mb2 time_exact32_c128_mb2 time_exact32_c160_mb2 time_exact64_c192_mb2 time_fast32 time_fast32_c160_mb2; do
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
done
cat $OUT
timeout 600 python bench.py --no-cpu-baseline --no-train > gpurun_out/r2c5_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2c5_bench.log
python - <<'PY'
import json
ls=[l for l in open('gpurun_out/r2c5_bench.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'fast',d.get('other_numerics',{}).get('value'))
    for k in d['roofline']['kernels']: print(k['layer'], round(k['us'],1), round(k['tflops'],1))
else:
    print(open('gpurun_out/r2c5_bench.log').read()[-1500:])
PY

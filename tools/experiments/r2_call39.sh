#!/bin/bash
# Round 2, call 39: 16-output variant of the exact dx kernel: parity, layer time vs padded-to-32, configs 3 / 5
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -k "16_outputs or dx_kernel_vs_per_tap" > gpurun_out/r2c39_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c39_pytest.log
grep -E "passed|failed|FAILED|rc=|Error|outside" gpurun_out/r2c39_pytest.log | head
OUT=gpurun_out/r2c39_nout16.log; : > $OUT
for c in exact16_c16 exact16_c64 exact16_c32_small time_exact16_c16_256 time_exact32_c16_256; do
  echo "== $c" >> $OUT
  timeout 120 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"|max_abs_err|rror' | cut -c1-260 >> $OUT
done
cat $OUT
timeout 1200 python -m pytest tests/test_head_gpu.py -m gpu -q > gpurun_out/r2c39_head.log 2>&1; echo "head pytest rc=$?" >> gpurun_out/r2c39_head.log
grep -E "passed|failed|FAILED|rc=|Error|outside" gpurun_out/r2c39_head.log | head
for t16 in 0 1; do
  BHSR_HEAD_TC16=$t16 timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c39_bench_$t16.log 2>&1
  python - <<PY
import json
ls=[l for l in open('gpurun_out/r2c39_bench_$t16.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d['train']
    print('tc16 $t16: train',round(t['value'],1),round(t['ms_per_step'],2),'eager',round(t.get('eager_ms_per_step',0),2),'loss',round(t['loss'],3),'clocks',d['clocks']['sm_mhz'])
else:
    print(open('gpurun_out/r2c39_bench_$t16.log').read()[-1500:])
PY
  BHSR_HEAD_TC16=$t16 timeout 900 python tools/bench_configs.py --config 5 --grids 2560 > gpurun_out/r2c39_cfg5_$t16.log 2>&1; echo "tc16 $t16 cfg5: $(tail -1 gpurun_out/r2c39_cfg5_$t16.log | cut -c90-200)"
done

#!/bin/bash
# Round 2, call 42: 16-output dx kernel with the deeper activation ring (two groups, 2 KB staging per warp -> 3 stages per plane)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -k "16_outputs" > gpurun_out/r2c42_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c42_pytest.log
grep -E "passed|failed|FAILED|rc=|Error|outside" gpurun_out/r2c42_pytest.log | head
OUT=gpurun_out/r2c42_ring.log; : > $OUT
for ns in 2 3; do
  echo "== time_exact16_c16_256 ASTAGES=$ns" >> $OUT
  BHSR_ASTAGES=$ns timeout 120 python tools/probe_conv_tc.py time_exact16_c16_256 0 2>&1 | grep -E '"ms"|rror' | cut -c1-200 >> $OUT
done
echo "== time_exact16_c16_256 auto" >> $OUT
timeout 120 python tools/probe_conv_tc.py time_exact16_c16_256 0 2>&1 | grep -E '"ms"|max_abs_err|rror' | cut -c1-200 >> $OUT
echo "== time_exact32_c16_256 (padded)" >> $OUT
timeout 120 python tools/probe_conv_tc.py time_exact32_c16_256 0 2>&1 | grep -E '"ms"|rror' | cut -c1-200 >> $OUT
cat $OUT

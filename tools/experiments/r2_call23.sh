#!/bin/bash
# Round 2, call 23: lean MMA issue (one wait / asm block / commit per phase) in the dxs kernel, fast numerics
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -k "dxs" > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c23_pytest.log
grep -E "passed|failed|FAILED|outside|rc=|Error" gpurun_out/r2c23_pytest.log | head -20
OUT=gpurun_out/r2c23_lean.log; : > $OUT
for cc in 64 96 128 160; do
  c2=time_fast32_c${cc}_mb2; [ $cc = 64 ] && c2=time_fast32
  echo "== $c2 old kernel" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c2 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
  for lean in 0 1; do
    echo "== $c2 DXS_MB=2 LEAN=$lean" >> $OUT
    BHSR_DXS_LEAN=$lean BHSR_DXS_MB=2 timeout 60 python tools/probe_conv_tc.py $c2 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
    for mb in 3 4; do
      echo "== time_fast32_c${cc}_mb$mb LEAN=$lean" >> $OUT
      BHSR_DXS_LEAN=$lean timeout 60 python tools/probe_conv_tc.py time_fast32_c${cc}_mb$mb 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
    done
  done
done
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so
for c in time_fast32 time_fast32_c160_mb2; do
  echo "== $c DXS_MB=2 LEAN timing" >> $OUT
  BHSR_DXS_MB=2 timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
done
echo "== time_fast32_c160_mb4 LEAN timing" >> $OUT
timeout 60 python tools/probe_conv_tc.py time_fast32_c160_mb4 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-520 >> $OUT
cat $OUT

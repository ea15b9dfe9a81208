#!/bin/bash
# Round 2, call 3: all gpu tests (no -x), and the instruction-fetch vs issue-slot diagnostic of the dx epilogue
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
tail -12 gpurun_out/r2c3_pytest.log
OUT=gpurun_out/r2c3_icache.log; : > $OUT
export BHSR_DEBUG_TIMING=1 BHSR_LIB=$PKG/lib/libbhsr_timing.so
for c in time_exact32_mb2 time_exact32_c160_mb2; do
  for m in 0 4 7 8; do
    echo "== $c timing NOMMA=$m" >> $OUT
    BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-420 >> $OUT
  done
done
cat $OUT

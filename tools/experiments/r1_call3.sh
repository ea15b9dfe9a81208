#!/bin/bash
# dx-in-N kernel bring-up: correctness probes, cycle counters, then the gpu tests and a bench line.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp7.log; : > $OUT
HERE=$(pwd)
for c in exact32 exact32_mb2 fast32 fast32_mb2 exact32_c96 exact32_c160_mb2 fast32_c160_mb2 fast32_c96_mb2 odd_h odd_h_mb2 small_multi small_multi_mb2 small_multi_exact; do
  echo "== $c" >> $OUT
  timeout 90 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror|bhsr:' | head -5 >> $OUT
  echo "rc=$?" >> $OUT
done
for c in exact32_mb2 exact32_c160_mb2; do
  echo "== $c FORCE_STREAM" >> $OUT
  BHSR_DEBUG_FORCE_STREAM=1 timeout 90 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror|bhsr:' | head -5 >> $OUT
done
export BHSR_DEBUG_TIMING=1
for c in time_exact32 time_exact32_mb2 time_exact32_c160_mb2 time_fast32 time_fast32_c160_mb2; do
  echo "== $c" >> $OUT
  BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
  echo "== $c NOMMA" >> $OUT
  BHSR_DEBUG_NOMMA=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
unset BHSR_DEBUG_TIMING
cat $OUT | cut -c1-600
timeout 900 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log; tail -8 gpurun_out/pytest_gpu7.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench7.log
tail -2 gpurun_out/bench7.log | cut -c1-1800

#!/bin/bash
# Round-1 re-entry call: gpu tests, bench line, in-kernel cycle counters, ncu launch list + full capture.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.get_device_name(0))" > gpurun_out/warm.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv > gpurun_out/smi_idle.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1500
OUT=gpurun_out/exp.log; : > $OUT
HERE=$(pwd)
for c in time_exact32 time_exact32_mb2 time_exact32_c160 time_exact32_c160_mb2 time_exact64_c192 time_exact64_c192_mb2 time_fast32 time_fast32_c160_mb2 time_fast64_c192; do
  echo "== $c" >> $OUT
  BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so BHSR_DEBUG_TIMING=1 timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
for c in time_exact32_mb2 time_exact64_c192_mb2; do
  echo "== $c NO_PDL" >> $OUT
  BHSR_NO_PDL=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so BHSR_DEBUG_TIMING=1 timeout 120 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' >> $OUT
done
cat $OUT | cut -c1-700
NUMERICS=exact bash tools/run_profile.sh

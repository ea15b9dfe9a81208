#!/bin/bash
# Round 2, call 1: sanity (gpu tests + bench with the round-1 product library) and the epilogue store-path
# diagnostics prepared at the end of round 1 (DESIGN.md §8 Finding 3): swizzled staging build, staging-only /
# stores-only / drain-only epilogues on the three dominant layer shapes.
mkdir -p gpurun_out
HERE=$(pwd)
PKG=$HERE/super-resolution-building-height-estimation_b200
python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))" > gpurun_out/r2c1_warm.log 2>&1
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,clocks.sm --format=csv >> gpurun_out/r2c1_warm.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -3 gpurun_out/r2c1_pytest.log
BHSR_LIB=$PKG/lib/libbhsr_episwz.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest_episwz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest_episwz.log
tail -3 gpurun_out/r2c1_pytest_episwz.log

OUT=gpurun_out/r2c1_epi.log; : > $OUT
for c in time_exact32_mb2 time_exact32_c96_mb2 time_exact32_c160_mb2 time_exact64_c192_mb2; do
  for lib in libbhsr.so libbhsr_episwz.so; do
    echo "== $c $lib" >> $OUT
    BHSR_LIB=$PKG/lib/$lib timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"' | cut -c1-200 >> $OUT
  done
done
export BHSR_DEBUG_TIMING=1
for c in time_exact32_mb2 time_exact32_c160_mb2; do
  for m in 0 4 5 6; do
    echo "== $c timing NOMMA=$m" >> $OUT
    BHSR_LIB=$PKG/lib/libbhsr_timing.so BHSR_DEBUG_NOMMA=$m timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-420 >> $OUT
  done
  echo "== $c timing_episwz NOMMA=0" >> $OUT
  BHSR_LIB=$PKG/lib/libbhsr_timing_episwz.so timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-420 >> $OUT
done
unset BHSR_DEBUG_TIMING
cat $OUT

timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c1_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2c1_bench.log
tail -2 gpurun_out/r2c1_bench.log | cut -c1-1200
BHSR_LIB=$PKG/lib/libbhsr_episwz.so timeout 600 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r2c1_bench_episwz.log 2>&1
tail -1 gpurun_out/r2c1_bench_episwz.log | cut -c1-600

#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp15.log; : > $OUT
export BHSR_PAIR=1
for c in exact64_c192_mb2 exact64_c64_mb2_nb6; do
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror|bhsr:|Traceback' | head -8 >> $OUT
  echo "rc=$?" >> $OUT
done
for c in time_exact64_c192_mb2; do
  echo "== $c pair" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"|max_abs_err|rror|bhsr:' | head -5 >> $OUT
  echo "== $c per-tap" >> $OUT
  BHSR_PAIR=0 timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"|max_abs_err' | head -5 >> $OUT
done
cat $OUT | cut -c1-400
if grep -q '"frac_bad": 0.0' $OUT; then
  timeout 600 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu15.log; tail -4 gpurun_out/pytest_gpu15.log
  timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-secondary > gpurun_out/bench15.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench15.log
  tail -2 gpurun_out/bench15.log | cut -c1-600
fi

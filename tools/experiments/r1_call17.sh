#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp17.log; : > $OUT
export BHSR_DX_PAIR=1
for c in exact32_mb2 exact32_c160_mb2_nb4 exact32_c160_mb2_nb6s exact32_c96_w130_nb2; do
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror|bhsr:|Traceback' | head -8 >> $OUT
  echo "rc=$?" >> $OUT
done
for c in time_exact32_mb2 time_exact32_c96_mb2 time_exact32_c128_mb2 time_exact32_c160_mb2; do
  echo "== $c pair" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"|max_abs_err|rror|bhsr:' | cut -c1-160 | head -5 >> $OUT
  echo "== $c single" >> $OUT
  BHSR_DX_PAIR=0 timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E '"ms"' | head -5 >> $OUT
done
cat $OUT | cut -c1-300
if [ $(grep -c '"frac_bad": 0.0' $OUT) -ge 8 ]; then
  timeout 600 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu17.log; tail -4 gpurun_out/pytest_gpu17.log
  timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-secondary > gpurun_out/bench17.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench17.log
  tail -2 gpurun_out/bench17.log | cut -c1-400
fi

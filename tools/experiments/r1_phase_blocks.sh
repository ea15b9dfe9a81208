#!/bin/bash
# One-block-per-phase issue in the dx kernel: correctness probes, cycle counters, gpu tests, bench.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp_phase.log; : > $OUT
HERE=$(pwd)
for c in exact32_mb2 exact32 fast32_mb2 exact32_c160_mb2_nb4 odd_h_mb2 exact32_c96_w130_nb2; do
  echo "== $c" >> $OUT
  timeout 60 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E 'max_abs_err|rror|bhsr:' | cut -c1-200 | head -4 >> $OUT
done
for c in time_exact32_mb2 time_exact32_c160_mb2 time_fast32; do
  echo "== $c" >> $OUT
  BHSR_DEBUG_TIMING=1 BHSR_LIB=$HERE/super-resolution-building-height-estimation_b200/lib/libbhsr_timing.so timeout 60 python tools/probe_conv_tc.py $c 0 2>/dev/null | grep -E '"ms"|cycles' | cut -c1-420 >> $OUT
done
cat $OUT
if [ $(grep -c '"frac_bad": 0.0' $OUT) -ge 6 ]; then
  timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  timeout 200 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_phase.log 2>&1
  tail -1 gpurun_out/bench_phase.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['other_numerics']['value'])"
fi

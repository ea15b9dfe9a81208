#!/bin/bash
# Run every probe case in its own process with a timeout; collect JSON lines.
mkdir -p gpurun_out
OUT=gpurun_out/probe_conv_tc.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT 2>&1
CASES=${CASES:-"fast32 fast32_mb2 exact32 fast64_c192 exact64_c192 exact32_c96 fast32_c160_mb2 up_exact up_fast_mb2 hr_exact_nchw hr_fast_nchw_mb2 odd_h"}
MODES=${MODES:-"0 1"}
for m in $MODES; do
  for c in $CASES; do
    echo "== $c mode $m" >> $OUT
    timeout 120 python tools/probe_conv_tc.py $c $m >> $OUT 2>&1
    echo "rc=$?" >> $OUT
  done
done
cat $OUT | grep -E "^==|^\{|rc=|Error|error|timeout" | head -150

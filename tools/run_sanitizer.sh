#!/bin/bash
# compute-sanitizer over small cases of the tensor-core conv kernels (memcheck / racecheck / synccheck).
# TOOLS and CASES select a subset.  Output: gpurun_out/sanitizer.log
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/sanitizer.log; : > $OUT
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  for c in ${CASES:-exact32_mb2 exact64_c192_mb2 exact32_c96_w130_nb2 fast32_mb2 odd_h_mb2}; do
    echo "== $tool $c ${EXTRA_ENV}" >> $OUT
    env $EXTRA_ENV timeout 240 compute-sanitizer --tool $tool --print-limit 5 python tools/probe_conv_tc.py $c 0 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max_abs_err|hazard|Invalid|Error" | cut -c1-240 | head -8 >> $OUT
  done
done
cat $OUT

#!/bin/bash
# Round 2, call 29: compute-sanitizer over the round-2 kernels: conv_dxs (probing and lean issuers, ragged tiles),
# float4 BatchNorm kernels + tensor-core head (one BasicBlock backward test), RRDBNet backward (one small case)
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
TOOLS="memcheck racecheck synccheck" CASES="fast32_mb4 fast32_c96_w130_mb3 fast32_c160_mb4" bash tools/run_sanitizer.sh > /dev/null 2>&1
cp gpurun_out/sanitizer.log gpurun_out/r2c29_sanitizer_dxs.log
TOOLS="memcheck racecheck" CASES="fast32_mb4 fast32_c96_w130_mb3" EXTRA_ENV="BHSR_DXS_LEAN=1" bash tools/run_sanitizer.sh > /dev/null 2>&1
cat gpurun_out/sanitizer.log >> gpurun_out/r2c29_sanitizer_dxs.log
cat gpurun_out/r2c29_sanitizer_dxs.log
OUT=gpurun_out/r2c29_sanitizer_tests.log; : > $OUT
for tool in memcheck racecheck; do
  echo "== $tool test_rrdbnet_backward (1 block, nb=3, 8x8)" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q -x -k "backward_vs_oracle and True-1-3-8" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard|Invalid" | cut -c1-240 | head -8 >> $OUT
  echo "== $tool test_hrfuse_residual_backward" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_head_gpu.py -m gpu -q -x -k "hrfuse_residual_backward" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard|Invalid" | cut -c1-240 | head -8 >> $OUT
done
cat $OUT

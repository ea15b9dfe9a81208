#!/bin/bash
# Last call of round 2: the whole gpu suite + smoke + both bench arms with the final library, and compute-sanitizer
# (memcheck, racecheck) over the fused CE + Dice kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.log 2>&1; tail -1 gpurun_out/bench_final_ref.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_final.log
tail -2 gpurun_out/bench_final.log | cut -c1-400
: > gpurun_out/sanitizer_ce_dice.log
for tool in memcheck racecheck; do
  echo "== $tool tools/sanitize_ce_dice.py" >> gpurun_out/sanitizer_ce_dice.log
  timeout 300 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_ce_dice.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ce_dice|hazard|Invalid|Error" | cut -c1-240 | head -12 >> gpurun_out/sanitizer_ce_dice.log
done
cat gpurun_out/sanitizer_ce_dice.log

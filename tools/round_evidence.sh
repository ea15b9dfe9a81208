#!/bin/bash
# Final round evidence: gpu tests, smoke, bench (both arms), launch list + ncu capture.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.log 2>&1; tail -1 gpurun_out/bench_final_ref.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_final.log
tail -2 gpurun_out/bench_final.log | cut -c1-1500
NUMERICS=exact bash tools/run_profile.sh

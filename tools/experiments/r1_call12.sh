#!/bin/bash
mkdir -p gpurun_out
NUMERICS=exact bash tools/run_profile.sh
timeout 600 python bench.py > gpurun_out/bench12.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench12.log
tail -2 gpurun_out/bench12.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench12_ref.log 2>&1; tail -1 gpurun_out/bench12_ref.log | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke12.log 2>&1; tail -1 gpurun_out/smoke12.log

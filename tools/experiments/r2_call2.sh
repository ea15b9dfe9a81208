#!/bin/bash
# Round 2, call 2: the new parity tests + bench with the train sub-record.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -15 gpurun_out/r2c2_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c2_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2c2_bench.log
tail -3 gpurun_out/r2c2_bench.log | cut -c1-2500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c2_bench_ref.log 2>&1
tail -1 gpurun_out/r2c2_bench_ref.log | cut -c1-600

#!/bin/bash
# Round-2: the new gpu tests against the staged UNMODIFIED reference (baseline/_ref) + the stand-alone block backward,
# and the reference arm of bench.py running the reference's own module on the box's host cores
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_live.py tests/test_rrdbnet_gpu.py -m gpu -q -s \
  -k "live_reference or standalone_blocks or error_behaviour" > gpurun_out/r2_live_reference_pytest.log 2>&1
echo "pytest rc=$?"
tail -25 gpurun_out/r2_live_reference_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_live_reference_arm.log 2>&1
echo "reference arm rc=$?"
tail -1 gpurun_out/r2_live_reference_arm.log | cut -c1-900
BHSR_CPU_ARM=port timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_port_arm.log 2>&1
tail -1 gpurun_out/r2_port_arm.log | cut -c1-300

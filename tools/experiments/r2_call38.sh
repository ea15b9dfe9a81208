#!/bin/bash
# Round 2, call 38: inference sweep (config 5) with the smp part on a forked stream / channels_last; N2 tests
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -q -k "predict or shard or postproc" > gpurun_out/r2c38_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c38_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c38_pytest.log | head
for nhwc in 0 1; do
  BHSR_SMP_NHWC=$nhwc timeout 900 python tools/bench_configs.py --config 5 --grids 2560 > gpurun_out/r2c38_cfg5_$nhwc.log 2>&1; echo "nhwc $nhwc: $(tail -1 gpurun_out/r2c38_cfg5_$nhwc.log | cut -c1-300)"
done

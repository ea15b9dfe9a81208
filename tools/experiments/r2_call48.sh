#!/bin/bash
# Round 2, call 48: conv_last of RRDBNet.forward on the tensor-core kernel: trunk tests, forward() time at B = 64, config 6
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1200 python -m pytest tests/test_rrdbnet_gpu.py -m gpu -q > gpurun_out/r2c48_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c48_pytest.log
grep -E "passed|failed|FAILED|rc=|Error|outside" gpurun_out/r2c48_pytest.log | head
python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import bhsr
from bhsr.rrdbnet import RRDBNet
torch.manual_seed(0)
net = RRDBNet(3, 3, scale=4, num_block=23).cuda().eval()
x = torch.rand(64, 3, 64, 64, device="cuda")
with torch.no_grad():
    for name, fn in (("forward_feature", net.forward_feature), ("forward", net.forward)):
        for _ in range(3): fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn(x)
        e1.record(); torch.cuda.synchronize()
        print(name, "B=64:", round(e0.elapsed_time(e1) / 5, 2), "ms")
PY
timeout 900 python tools/bench_configs.py --config 6 --steps 3 --warmup 1 > gpurun_out/r2c48_cfg6.log 2>&1; tail -1 gpurun_out/r2c48_cfg6.log | cut -c1-200

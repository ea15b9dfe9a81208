#!/bin/bash
# Round-2: fused CE + Dice loss kernel — parity tests, the live-reference loss test, the training-step tests, and the
# training step's time with / without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_head_gpu.py tests/test_reference_live.py -m gpu -q -s \
  -k "ce_dice or weighted_losses or full_pipeline_train_step or graphed_train_step or epoch_harness or head_training_step" \
  > gpurun_out/r2_ce_dice_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2_ce_dice_pytest.log
for f in 1 0; do
  BHSR_FUSED_CE_DICE=$f timeout 600 python bench.py --no-cpu-baseline --no-secondary --steps 10 > gpurun_out/r2_ce_dice_bench_$f.log 2>&1
  python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/r2_ce_dice_bench_{f}.log') if l.startswith('{')][-1])
    print('fused' if f=='1' else 'stock', 'train %.1f tiles/s %.3f ms loss %.4f  fwd %.1f'%(d['train']['value'],d['train']['ms_per_step'],d['train']['loss'],d['value']))
except Exception as e:
    print(f,'FAILED',e); print(open(f'gpurun_out/r2_ce_dice_bench_{f}.log').read()[-1500:])
PY
done

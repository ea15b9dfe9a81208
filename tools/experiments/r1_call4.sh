#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
OUT=gpurun_out/exp4.log; : > $OUT
for b in 16 24 32 48 64; do
  echo "== batch $b" >> $OUT
  timeout 300 python bench.py --steps 5 --batch $b --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])" >> $OUT
done
echo "== batch 64 mblocks1" >> $OUT
BHSR_MBLOCKS=1 timeout 300 python bench.py --steps 5 --batch 64 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" >> $OUT
cat $OUT

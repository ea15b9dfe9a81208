#!/bin/bash
# Round-2 profile evidence: launch list of one forward step + ncu --set full capture of 10 consecutive trunk convs
mkdir -p gpurun_out
NUMERICS=exact bash tools/run_profile.sh
timeout 900 python bench.py --no-cpu-baseline --no-train > gpurun_out/r2e1_bench.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e1_bench.log') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'fast',d.get('other_numerics',{}).get('value'))
PY

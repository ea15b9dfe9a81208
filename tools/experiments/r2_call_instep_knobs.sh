#!/bin/bash
# Round-2: (1) the pipelined full-output e2e leg; (2) the opt-in dx-kernel variants measured IN the power-capped step
# (their round-1/2 verdicts came from 50-launch bursts at burst clocks; the whole step runs at the 600 W cap, where
# fewer shared-memory / L2 bytes per MAC can buy clock)
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-train --no-cpu-baseline --no-secondary --steps 20 > gpurun_out/r2k_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/r2k_{n}.log') if l.startswith('{')][-1])
    print(n,'value %.1f tiles/s  %.3f ms  e2e %.1f  clocks %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks']['sm_mhz']), [ (k['layer'],round(k['us'],1)) for k in d['roofline']['kernels']])
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r2k_{n}.log').read()[-1500:])
PY
}
timeout 600 python bench.py --no-train --no-cpu-baseline --steps 10 > gpurun_out/r2k_full.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2k_full.log') if l.startswith('{')][-1])
print('default value %.1f e2e %.1f full %s'%(d['value'],d['e2e']['value'],json.dumps(d.get('e2e_full_output_d2h'))[:400]))
PY
run default BHSR_NOOP=1
run dxpair BHSR_DX_PAIR=1
run dxlean BHSR_DX_LEAN=1
run dxpair_lean BHSR_DX_PAIR=1 BHSR_DX_LEAN=1
run default2 BHSR_NOOP=1

#!/bin/bash
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
timeout 1200 python -m pytest tests/test_head_gpu.py -m gpu -q -s > gpurun_out/r2c11_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c11_head.log
grep -E "passed|failed|FAILED|outside|rel-L2|rc=|illegal|Error" gpurun_out/r2c11_head.log | head -30
timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 5 --warmup 3 > gpurun_out/r2c11_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2c11_bench.log
python - <<'PY'
import json
ls=[l for l in open('gpurun_out/r2c11_bench.log') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); t=d.get('train',{})
    print('fwd',d['value'],'train',t.get('value'),t.get('ms_per_step'),'eager',t.get('eager_ms_per_step'),t.get('launch'),'loss',t.get('loss'))
else:
    print(open('gpurun_out/r2c11_bench.log').read()[-2500:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 9000 --csv --log-file gpurun_out/r2c11_launches_cfg3.csv python tools/bench_configs.py --config 3 --steps 2 --warmup 2 > gpurun_out/r2c11_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows=list(csv.reader(open('gpurun_out/r2c11_launches_cfg3.csv')))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r)); name=re.sub(r'\(.*','',d.get('Kernel Name',''))[:80]
    try: v=float(d.get('Metric Value','0').replace(',',''))
    except: continue
    u=d.get('Metric Unit','')
    if u=='ns': v/=1000
    elif u=='ms': v*=1000
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values()); print('total us',round(tot),'launches',sum(v[0] for v in agg.values()))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]: print(f"{v[1]:9.0f} us {v[0]:5d} {100*v[1]/tot:5.1f}%  {k}")
PY

#!/bin/bash
# forward pass of the round-2 loop: GPU tests, ncu launch list of the nearest-neighbour kernels, quick bench, ncu launch list of
# one eager fwd+loss step
mkdir -p gpurun_out
TAG=${1:-r02l}
for f in test_gpu_ops test_gpu_detector test_gpu_vs_reference; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest_${f}_$TAG.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/pytest_${f}_$TAG.log)"
done
grep -h "AssertionError\|Error" gpurun_out/pytest_test_gpu_*_$TAG.log | head -10 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nn_launches_$TAG.csv python tools/nn_microbench.py > gpurun_out/nn_micro_ncu_$TAG.json 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/nn_launches_$TAG.csv') if not l.startswith('=='))]
hdr=rows[0]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    if len(r)<=mv or 'usip' not in r[kn]: continue
    k=r[kn].split('(')[0][-40:]
    try: v=float(r[mv].replace(',',''))
    except: continue
    agg.setdefault(k,[]).append(v)
for k,v in agg.items():
    v=sorted(v); print("%-42s n=%3d median %.1f us min %.1f" % (k, len(v), v[len(v)//2]/1e3, v[0]/1e3))
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-descriptor --no-tf32-backward > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value']); print(j.get('train_step'))
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --profile-from-start off --log-file gpurun_out/fwd_launches_$TAG.csv python tools/ncu_step.py fwd > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/fwd_launches_$TAG.csv') if not l.startswith('=='))]
hdr=rows[0]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    if len(r)<=mv: continue
    k=r[kn].split('(')[0][-44:]
    try: v=float(r[mv].replace(',',''))
    except: continue
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print('fwd launch list total us', tot/1e3)
for k,(n,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:40]: print("%-46s n=%3d total %.1f us" % (k,n,v/1e3))
PY

#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02m}
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "pairwise_min or som_assign" > gpurun_out/pytest_nn_$TAG.log 2>&1
echo "== nn tests: $(tail -1 gpurun_out/pytest_nn_$TAG.log)"
timeout 900 python -m pytest tests/test_gpu_detector.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_test_gpu_detector_$TAG.log 2>&1
echo "== detector: $(tail -1 gpurun_out/pytest_test_gpu_detector_$TAG.log)"
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

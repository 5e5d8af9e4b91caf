#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pca_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/eig_bench.py 2>&1 | tail -5
EIG_N=2504 EIG_REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:tridiag -s 2600 -c 3000 --csv --log-file gpurun_out/launches_eig.csv python tools/eig_bench.py > gpurun_out/ncu_eig.log 2>&1
tail -2 gpurun_out/ncu_eig.log
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_eig.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
d=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>=15: d[r[4].split('(')[0][-30:]].append(float(r[-1]))
for k,v in d.items():
    q=len(v)//4
    print(k, len(v), 'avg us %.2f'%(sum(v)/len(v)/1e3), 'first-quarter avg %.2f'%(sum(v[:q])/q/1e3), 'last-quarter avg %.2f'%(sum(v[-q:])/q/1e3))
PY

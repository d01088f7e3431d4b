#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4500 -c 4300 --csv --log-file gpurun_out/r66_train_launches.csv python bench.py --workload train --steps 1 --warmup 1 > gpurun_out/r66_log.txt 2>&1
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/r66_train_launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.defaultdict(lambda:[0,0.0])
for x in rows:
    n=x['Kernel Name'].split('(')[0][-40:]; a=agg[n]; a[0]+=1; a[1]+=float(x['Metric Value'].replace(',',''))
tot=sum(a[1] for a in agg.values())
print(len(rows),'launches, total ms',tot/1e6)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{k:42s} n={a[0]:5d} {a[1]/1e6:8.2f} ms {a[1]/tot:.3f}")
PY

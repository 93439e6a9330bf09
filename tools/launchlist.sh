#!/bin/bash
# ncu launch list of the bench step (graph mode); summarised per kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(list)
for row in csv.DictReader(lines):
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    agg[row['Kernel Name'][:80]].append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:82s} n={len(v):4d} total={sum(v):8.1f}us mean={sum(v)/len(v):6.2f} min={min(v):6.2f}")
PY

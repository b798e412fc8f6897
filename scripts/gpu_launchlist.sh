#!/bin/bash
# warm-cache launch list (no cache flush between kernels): per-kernel durations of eager steps
set -u
mkdir -p gpurun_out
TAG=${1:-default}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv \
  --log-file gpurun_out/launches_warm_$TAG.csv python bench.py --model ${MODEL:-deepfm} --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list $TAG exit $?"
python - <<PY
import csv
with open('gpurun_out/launches_warm_$TAG.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
rows=[(int(x['ID']), x['Kernel Name'][:56], int(x['Metric Value']), x['Grid Size']) for x in csv.DictReader(lines)]
idx=[i for i,r in enumerate(rows) if 'criteo_rows' in r[1]]
s=idx[5]; e=idx[6]
tot=0
for r in rows[s-1:e-1]:
    print(r); tot+=r[2]
print('sum', tot)
PY

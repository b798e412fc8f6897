#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for M in din xdeepfm; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
  --log-file gpurun_out/launches_${M}.csv python bench.py --model $M --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 2 > gpurun_out/ncu_bench_${M}.log 2>&1
echo "$M launch list exit $?"; wc -l gpurun_out/launches_${M}.csv
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cin_tc_kernel|din_att" -s 4 -c 6 \
  -f -o gpurun_out/prof_cin python bench.py --model xdeepfm --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 2 > gpurun_out/ncu_full_cin.log 2>&1
echo "cin full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"din_att" -s 2 -c 4 \
  -f -o gpurun_out/prof_din python bench.py --model din --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 2 > gpurun_out/ncu_full_din.log 2>&1
echo "din full exit $?"; ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# ncu evidence: per-launch durations of one eager step + full capture of the hot kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MODEL=${MODEL:-deepfm}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-700} --csv \
  --log-file gpurun_out/launches_${MODEL}.csv python bench.py --model $MODEL --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 ${BENCH_EXTRA:-} > gpurun_out/ncu_bench_${MODEL}.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_${MODEL}.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-embed_|adam_rows}" -s ${NCU_SKIP:-12} -c ${NCU_FULL_COUNT:-6} \
  -f -o gpurun_out/prof_${MODEL} python bench.py --model $MODEL --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 ${BENCH_EXTRA:-} > gpurun_out/ncu_full_${MODEL}.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# round 2, call C: ncu --set full on the fused scatter+Adam kernel and the count kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on \
  -k regex:"embed_bwd_adam|count_rows" -s 18 -c 6 -o gpurun_out/r02c_prof_bwd_adam -f \
  python bench.py --model deepfm --steps 8 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r02c_ncu.log 2>&1
tail -5 gpurun_out/r02c_ncu.log
ls -la gpurun_out/*.ncu-rep

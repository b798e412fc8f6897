#!/bin/bash
# Bench every model family once (N=1) -> gpurun_out/bench_<model>.json
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for M in ${MODELS:-fm dcn xdeepfm din}; do
  timeout 600 python bench.py --model $M --steps ${BENCH_STEPS:-100} --warmup 5 ${EXTRA:-} > gpurun_out/bench_$M.json 2> gpurun_out/bench_$M.err
  echo "== $M exit $?"; grep -v Warning gpurun_out/bench_$M.err | tail -4; cut -c1-1800 gpurun_out/bench_$M.json
done

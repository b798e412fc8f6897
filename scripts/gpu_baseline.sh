#!/bin/bash
# One gpurun call: full GPU test suite, smoke, deepfm bench, launch list, full ncu capture of the embed kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu --timeout 240 --timeout-method=thread --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - T0 ))s" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|Error:|error:" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -20
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
echo "== bench at $(( $(date +%s) - T0 ))s"
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_deepfm.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_deepfm.json
echo "== launch list at $(( $(date +%s) - T0 ))s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_deepfm.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_deepfm.csv
echo "== full capture at $(( $(date +%s) - T0 ))s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"embed_|adam_rows" -s 12 -c 6 \
  -f -o gpurun_out/prof_embed python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $? at $(( $(date +%s) - T0 ))s"; ls -la gpurun_out/*.ncu-rep

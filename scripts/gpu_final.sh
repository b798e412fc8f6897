#!/bin/bash
# Round-end evidence in one gpurun call: tests, smoke, bench of every config (+ the reference arm),
# launch lists, full ncu capture of the hot kernels, in-graph kernel timeline.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu --timeout 240 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - T0 ))s"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
for m in deepfm fm dcn xdeepfm din; do
  timeout 600 python bench.py --model $m --steps ${BENCH_STEPS:-200} --warmup 5 > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "bench $m exit $? at $(( $(date +%s) - T0 ))s"; cut -c1-240 gpurun_out/bench_$m.json
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>/dev/null; echo "reference exit $?"; cut -c1-200 gpurun_out/bench_reference.json
echo "== launch lists at $(( $(date +%s) - T0 ))s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_deepfm.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
echo "== full capture at $(( $(date +%s) - T0 ))s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"embed_|adam_rows|tower_mid|tc_gemm" -s 21 -c 8 \
  -f -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $? at $(( $(date +%s) - T0 ))s"
for m in deepfm fm xdeepfm; do python scripts/trace_step.py --model $m > gpurun_out/trace_$m.txt 2>&1; done
head -20 gpurun_out/trace_deepfm.txt
echo "done at $(( $(date +%s) - T0 ))s"

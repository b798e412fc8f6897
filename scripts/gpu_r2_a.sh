#!/bin/bash
# round 2, call A: GPU tests + smoke + DeepFM / DCN / DIN bench lines.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|Error:|error:|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -40
grep -E "config3" gpurun_out/pytest_gpu.log | tail
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
for m in deepfm dcn din; do
  echo "== bench $m"
  timeout 600 python bench.py --model $m --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_$m.json 2> gpurun_out/r02a_bench_$m.err; echo "bench exit $?"; tail -3 gpurun_out/r02a_bench_$m.err; cut -c1-1500 gpurun_out/r02a_bench_$m.json
done

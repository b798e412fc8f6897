#!/bin/bash
# One gpurun call: GPU tests, smoke, bench.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -n "${DEBUG_SCRIPT:-}" ]; then timeout 300 python $DEBUG_SCRIPT > gpurun_out/debug.log 2>&1; tail -40 gpurun_out/debug.log; fi
echo "== pytest -m gpu (everything but the tcgen05 CIN)"
timeout 1200 python -m pytest tests -q -m gpu -k "not tf32" --timeout 240 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu.log | tail -40
echo "== pytest -m gpu (tcgen05 CIN)"
timeout 600 python -m pytest tests -q -m gpu -k "tf32" --timeout 120 --timeout-method=thread > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" | tee -a gpurun_out/pytest_tc.log
grep -E "passed|failed|^FAILED|Error:|error:|^E  .*err" gpurun_out/pytest_tc.log | cut -c1-400 | tail -40
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-200} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json

#!/bin/bash
# round 2, call B: GPU tests + DeepFM bench with the fused scatter+Adam, in-graph trace
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|Error:|error:|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -40
for m in deepfm fm; do
  echo "== bench $m"
  timeout 600 python bench.py --model $m --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_$m.json 2> gpurun_out/r02b_bench_$m.err; echo "bench exit $?"; tail -3 gpurun_out/r02b_bench_$m.err; cut -c1-300 gpurun_out/r02b_bench_$m.json
done
echo "== bench deepfm unfused"
CTR_FUSED_ROW_ADAM=0 timeout 600 python bench.py --model deepfm --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_deepfm_unfused.json 2> gpurun_out/r02b_bench_deepfm_unfused.err; cut -c1-300 gpurun_out/r02b_bench_deepfm_unfused.json
echo "== trace"
timeout 300 python scripts/trace_step.py --model deepfm > gpurun_out/r02b_trace_deepfm.txt 2>&1; tail -40 gpurun_out/r02b_trace_deepfm.txt

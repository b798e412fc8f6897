#!/bin/bash
# round 2, call E: full GPU tests, xDeepFM bench (tf32x3 + tf32) with trace, DeepFM bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|Error:|error:|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -30
for prec in tf32x3 tf32; do
  echo "== bench xdeepfm $prec"
  timeout 900 python bench.py --model xdeepfm --cin-precision $prec --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_xdeepfm_$prec.json 2> gpurun_out/r02e_bench_xdeepfm_$prec.err; echo "bench exit $?"; tail -3 gpurun_out/r02e_bench_xdeepfm_$prec.err; cut -c1-400 gpurun_out/r02e_bench_xdeepfm_$prec.json
done
echo "== trace xdeepfm (tf32x3)"
timeout 300 python scripts/trace_step.py --model xdeepfm > gpurun_out/r02e_trace_xdeepfm.txt 2>&1; grep -E "steps, span|cin_|ctr::" gpurun_out/r02e_trace_xdeepfm.txt | cut -c1-120 | head -40
echo "== bench deepfm"
timeout 600 python bench.py --model deepfm --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_deepfm.json 2> gpurun_out/r02e_bench_deepfm.err; echo "bench exit $?"; tail -3 gpurun_out/r02e_bench_deepfm.err; cut -c1-300 gpurun_out/r02e_bench_deepfm.json

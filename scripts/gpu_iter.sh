#!/bin/bash
# One gpurun call per iteration: full GPU tests, deepfm bench (graph), launch list of an eager step.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -x --timeout 240 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - T0 ))s"
grep -E "passed|failed|^FAILED|^E  |Error:|error:" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -25
for m in ${BENCH_MODELS:-deepfm}; do
  echo "== bench $m at $(( $(date +%s) - T0 ))s"
  timeout 600 python bench.py --model $m --steps ${BENCH_STEPS:-200} --warmup 5 ${BENCH_EXTRA:-} > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  rc=$?; echo "bench exit $rc"; tail -3 gpurun_out/bench_$m.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("$m value %.4g  ms/step %.4f  e2e %.4g  launches/step %s  roofline frac %.3f fwd %s bwd %s adam %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step"), r.get("frac", 0),
        r.get("fwd", {}).get("us"), r.get("bwd", {}).get("us"), r.get("adam_rows_us")))
except Exception as e:
    print("no bench line:", e)
PY
  if [ $rc -ne 0 ]; then
    echo "== retry with CTR_MID_COOP=0"
    CTR_MID_COOP=0 timeout 600 python bench.py --model $m --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${m}_nocoop.json 2> gpurun_out/bench_${m}_nocoop.err
    echo "exit $?"; tail -3 gpurun_out/bench_${m}_nocoop.err; cat gpurun_out/bench_${m}_nocoop.json | cut -c1-400
  fi
done
if [ -n "${LAUNCH_LIST:-1}" ]; then
  echo "== launch list at $(( $(date +%s) - T0 ))s"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_deepfm.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --n-batches 4 > gpurun_out/ncu_bench.log 2>&1
  echo "launch list exit $?"; wc -l gpurun_out/launches_deepfm.csv
fi
echo "done at $(( $(date +%s) - T0 ))s"

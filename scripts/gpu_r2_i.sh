#!/bin/bash
# round 2, iteration call (1 GPU): selected tests, deepfm bench + timeline; env TESTS / EXTRA
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${TAG:-r02i}
echo "== pytest ${TESTS:-tests}"
timeout 1200 python -m pytest ${TESTS:-tests} -q -m gpu -x --timeout 300 --timeout-method=thread ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^E  |Error:|error:" gpurun_out/${TAG}_pytest.log | cut -c1-300 | tail -15
for m in ${BENCH_MODELS:-deepfm}; do
  timeout 600 python bench.py --model $m --steps 200 --warmup 5 ${BENCH_EXTRA:---no-cpu-baseline} > gpurun_out/${TAG}_bench_$m.json 2> gpurun_out/${TAG}_bench_$m.err
  echo "bench $m exit $?"; tail -2 gpurun_out/${TAG}_bench_$m.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$m.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("$m value %.4g  ms/step %.4f  e2e %.4g  launches/step %s  roofline frac %.3f fwd %s bwd %s adam %s large %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step"), r.get("frac", 0),
        r.get("fwd", {}).get("us"), r.get("bwd", {}).get("us"), r.get("adam_rows_us"), r.get("large_batch")))
except Exception as e:
    print("no bench line:", e)
PY
  timeout 300 python scripts/trace_step.py --model $m --timeline > gpurun_out/${TAG}_trace_$m.txt 2>&1
  grep -v Warning gpurun_out/${TAG}_trace_$m.txt | tail -60
done
if [ -n "${EXTRA_CMD:-}" ]; then
  echo "== extra: $EXTRA_CMD"
  bash -c "$EXTRA_CMD"
fi

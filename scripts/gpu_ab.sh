#!/bin/bash
# A/B runs of the deepfm bench under environment settings: each line of $AB is "tag|ENV=.. ENV=.."
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "$AB" | while IFS='|' read -r tag envs; do
  [ -z "$tag" ] && continue
  env $envs timeout 300 python bench.py --model ${MODEL:-deepfm} --steps 300 --warmup 10 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("%-28s ms/step %.4f  e2e ms %.4f  fwd %.2f bwd %.2f adam %.2f" % ("$tag", d["ms_per_step"], d["e2e"]["ms_per_step"],
        r.get("fwd", {}).get("us", 0), r.get("bwd", {}).get("us", 0), r.get("adam_rows_us", 0)))
except Exception as e:
    print("$tag: no bench line:", e); print(open("gpurun_out/ab_$tag.err").read()[-1500:])
PY
done
if [ -n "${TRACE_ENV:-}" ]; then
  env $TRACE_ENV timeout 300 python scripts/trace_step.py --model ${MODEL:-deepfm} --timeline 2>&1 | grep -v -i warn | tail -40
fi

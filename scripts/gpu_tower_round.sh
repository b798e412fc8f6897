#!/bin/bash
# tower tests + microbench + model tests + bench in one call
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONPATH=.
timeout 300 python -m pytest tests/test_gpu_tower.py -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_tower.py > gpurun_out/bench_tower.log 2>&1; cat gpurun_out/bench_tower.log
timeout 600 python -m pytest tests/test_gpu_models.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"

#!/bin/bash
# round 2, call H (1 GPU): validate the restructured scatter-add (ctr_embed_bwd v2)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/r02h_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02h_pytest_gpu.log | cut -c1-250 | tail -12
b() { tag=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r02h_bench_$tag.json 2> gpurun_out/r02h_bench_$tag.err; echo "bench $tag exit $?: $(grep '^{' gpurun_out/r02h_bench_$tag.json | cut -c1-200)"; }
b deepfm --model deepfm --steps 200 --warmup 5 --no-cpu-baseline
b deepfm_zipf --model deepfm --dist zipf --steps 200 --warmup 5 --no-cpu-baseline
b deepfm_zipf_noagg --model deepfm --dist zipf --no-bwd-aggregate --steps 200 --warmup 5 --no-cpu-baseline
b fm --model fm --steps 200 --warmup 5 --no-cpu-baseline
b dcn --model dcn --steps 200 --warmup 5 --no-cpu-baseline
timeout 300 python scripts/trace_step.py --model deepfm > gpurun_out/r02h_trace_deepfm.txt 2>&1; tail -16 gpurun_out/r02h_trace_deepfm.txt | cut -c1-110
timeout 300 python scripts/trace_step.py --model din > gpurun_out/r02h_trace_din.txt 2>&1; tail -12 gpurun_out/r02h_trace_din.txt | cut -c1-140

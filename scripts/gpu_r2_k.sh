#!/bin/bash
# round 2, call K (1 GPU): evidence run after the fused lookup + first-layer kernel - tests, every
# bench line, traces, launch list, ncu of the step's kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${TAG:-r02k}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest_gpu.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?: $(tail -1 gpurun_out/${T}_smoke.log | cut -c1-160)"
b() {  # tag, args...   (BENV: extra environment for this line)
  tag=$1; shift
  env ${BENV:-A=1} timeout 900 python bench.py "$@" > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err
  echo "bench $tag exit $?: $(grep '^{' gpurun_out/${T}_bench_$tag.json | cut -c1-230)"
}
b deepfm --model deepfm --steps 200 --warmup 5
b reference --impl reference --steps 30 --warmup 3
if [ -z "${QUICK:-}" ]; then
BENV="CTR_FUSED_L0=0" b deepfm_unfused --model deepfm --steps 200 --warmup 5 --no-cpu-baseline
BENV="CTR_FUSED_L0=0 CTR_PREFETCH_IDS=0 CTR_DENSE_ON_SIDE=0 CTR_GRAPH_DOUBLE=0" b deepfm_round_start --model deepfm --steps 200 --warmup 5 --no-cpu-baseline
b deepfm_zipf --model deepfm --dist zipf --steps 200 --warmup 5 --no-cpu-baseline
b deepfm_exact_tf_ref --model deepfm --table ref --embedding-adam exact_tf --steps 100 --warmup 5 --no-cpu-baseline
b fm --model fm --steps 200 --warmup 5 --cpu-seconds 8
b dcn --model dcn --steps 200 --warmup 5 --cpu-seconds 8
b din --model din --steps 100 --warmup 5
b xdeepfm_tf32x3 --model xdeepfm --cin-precision tf32x3 --steps 30 --warmup 3 --cpu-seconds 10
b xdeepfm_tf32 --model xdeepfm --cin-precision tf32 --steps 30 --warmup 3 --no-cpu-baseline
fi
for m in ${TRACE_MODELS:-deepfm dcn din xdeepfm}; do
  timeout 300 python scripts/trace_step.py --model $m --timeline > gpurun_out/${T}_trace_$m.txt 2>&1; echo "trace $m: $(grep 'steps, span' gpurun_out/${T}_trace_$m.txt)"
done
echo "== ncu launch list (deepfm, eager steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${T}_launches_deepfm.csv python bench.py --model deepfm --steps 6 --warmup 3 --no-cpu-baseline --eager > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu --set full: the step's embedding kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"embed_tower_fwd_kernel|embed_bwd_kernel|adam_rows_kernel|tower_mid_kernel" -s 12 -c 8 -o gpurun_out/${T}_prof_step -f python bench.py --model deepfm --steps 6 --warmup 3 --no-cpu-baseline --eager > gpurun_out/${T}_ncu_step.log 2>&1; echo "ncu step exit $?"
ls -la gpurun_out/*.ncu-rep | tail -3

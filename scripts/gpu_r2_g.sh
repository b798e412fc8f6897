#!/bin/bash
# round 2, call G (1 GPU): the evidence run - tests, every bench line, traces, ncu
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/r02g_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r02g_pytest_gpu.log | cut -c1-200
b() {  # tag, args...
  tag=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/r02g_bench_$tag.json 2> gpurun_out/r02g_bench_$tag.err
  echo "bench $tag exit $?: $(grep '^{' gpurun_out/r02g_bench_$tag.json | cut -c1-200)"
}
b deepfm --model deepfm --steps 200 --warmup 5
b deepfm_zipf --model deepfm --dist zipf --steps 200 --warmup 5 --no-cpu-baseline
b deepfm_zipf_noagg --model deepfm --dist zipf --no-bwd-aggregate --steps 200 --warmup 5 --no-cpu-baseline
b deepfm_exact_tf_ref --model deepfm --table ref --embedding-adam exact_tf --steps 100 --warmup 5 --no-cpu-baseline
b deepfm_lazy_ref --model deepfm --table ref --steps 200 --warmup 5 --no-cpu-baseline
b fm --model fm --steps 200 --warmup 5 --cpu-seconds 8
b dcn --model dcn --steps 200 --warmup 5 --cpu-seconds 8
b din --model din --steps 100 --warmup 5
b xdeepfm_tf32x3 --model xdeepfm --cin-precision tf32x3 --steps 30 --warmup 3 --cpu-seconds 10
b xdeepfm_tf32 --model xdeepfm --cin-precision tf32 --steps 30 --warmup 3 --no-cpu-baseline
b reference --impl reference --steps 30 --warmup 3
for m in deepfm dcn din xdeepfm; do
  timeout 300 python scripts/trace_step.py --model $m > gpurun_out/r02g_trace_$m.txt 2>&1; echo "trace $m: $(grep 'steps, span' gpurun_out/r02g_trace_$m.txt)"
done
echo "== ncu launch list (deepfm, eager steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02g_launches_deepfm.csv python bench.py --model deepfm --steps 6 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r02g_ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu --set full: embed fwd / bwd / adam (uniform)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"embed_fwd_kernel|embed_bwd_kernel|adam_rows_kernel" -s 12 -c 6 -o gpurun_out/r02g_prof_embed -f python bench.py --model deepfm --steps 6 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r02g_ncu_embed.log 2>&1; echo "ncu embed exit $?"
echo "== ncu --set full: embed bwd (zipf)"
timeout 600 ncu --set full --clock-control none -k regex:"embed_bwd_kernel" -s 6 -c 2 -o gpurun_out/r02g_prof_embed_bwd_zipf -f python bench.py --model deepfm --dist zipf --steps 6 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r02g_ncu_zipf.log 2>&1; echo "ncu zipf exit $?"
echo "== ncu --set full: CIN kernels"
timeout 900 ncu --set full --clock-control none -k regex:"cin_tc_kernel|cin_dw_fused_kernel" -s 10 -c 8 -o gpurun_out/r02g_prof_cin -f python bench.py --model xdeepfm --steps 3 --warmup 3 --no-cpu-baseline --eager > gpurun_out/r02g_ncu_cin.log 2>&1; echo "ncu cin exit $?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# round 2, call N (1 GPU): the DeepFM step under each of this round's switches, same box, back to back
# (each line: bench.py --steps 300 --warmup 10, value = HBM-resident, e2e = pinned host blob in, loss out)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r02n_ab_deepfm_step_variants.txt
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --model deepfm --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/r02n_$tag.json 2> gpurun_out/r02n_$tag.err
  python - "$tag" "$*" <<'PY' >> $OUT
import json, sys
tag, envs = sys.argv[1], sys.argv[2]
try:
    d = json.loads([l for l in open("gpurun_out/r02n_%s.json" % tag) if l.startswith("{")][-1])
    print("%-34s ms/step %.4f  (%.2f M samples/s)   e2e ms/step %.4f   [%s]" % (
        tag, d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], envs))
except Exception as e:
    print("%-34s failed: %r" % (tag, e))
PY
}
run default A=1
run default_again A=1
run lookup_unfused CTR_FUSED_L0=0
run no_id_prefetch CTR_PREFETCH_IDS=0
run dense_adam_on_main_stream CTR_DENSE_ON_SIDE=0
run single_graph_with_staging_copy CTR_GRAPH_DOUBLE=0
run fused_data_gradient_scatter CTR_FUSED_BWD0=1 CTR_DW_DEFER=0
run field_major_row_adam CTR_ADAM_ROWS_BF=1
run record_prefetch_in_lookup CTR_FUSED_L0=0 CTR_PREFETCH_IDS=0 CTR_OPTIONS=fwd_prefetch_record=1
run round_start_config CTR_FUSED_L0=0 CTR_PREFETCH_IDS=0 CTR_DENSE_ON_SIDE=0 CTR_GRAPH_DOUBLE=0
run default_third A=1
cat $OUT
grep '^{' gpurun_out/r02n_default_third.json | tail -1 > gpurun_out/r02n_bench_deepfm.json

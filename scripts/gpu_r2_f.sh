#!/bin/bash
# round 2, call F (8 GPUs): exchange parity at 8 ranks, then the config-5 bench (1e9-row table) on the
# peer-memory exchange (uniform + zipf) and on NCCL, with in-graph traces
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu8.txt 2>&1
echo "== parity worker (8 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/p2p_worker.py > gpurun_out/r02f_parity_n8.log 2>&1
echo "worker exit $?"; grep -E "PARITY|Error|error|Traceback" gpurun_out/r02f_parity_n8.log | cut -c1-500 | tail -8
run() {  # tag exchange dist
  echo "== bench --gpus 8 ($1)"
  CTR_SHARD_EXCHANGE=$2 CTR_TRACE=gpurun_out/r02f_trace_n8_$1.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 5 --dist $3 --no-cpu-baseline > gpurun_out/r02f_bench_n8_$1.json 2> gpurun_out/r02f_bench_n8_$1.err
  echo "bench exit $?"; grep -E "Error|error|Traceback|overflow" gpurun_out/r02f_bench_n8_$1.err | tail -5 | cut -c1-300; grep "^{" gpurun_out/r02f_bench_n8_$1.json | cut -c1-260
}
run p2p p2p uniform
run p2p_zipf p2p zipf
run nccl nccl uniform

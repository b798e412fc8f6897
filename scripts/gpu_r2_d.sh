#!/bin/bash
# round 2, call D (2 GPUs): sharded exchange parity on hardware + a short 2-GPU bench with trace
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu2.txt 2>&1
echo "== parity worker (p2p + nccl)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/p2p_worker.py > gpurun_out/r02d_parity.log 2>&1
echo "worker exit $?"; grep -E "PARITY|Error|error|Traceback" gpurun_out/r02d_parity.log | cut -c1-600 | tail -12
echo "== pytest sharded (skipped in this call)"; if false; then
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 600 --timeout-method=thread > gpurun_out/r02d_pytest_sharded.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r02d_pytest_sharded.log | cut -c1-400
fi
for ex in p2p; do
  echo "== bench --gpus 2 ($ex)"
  CTR_SHARD_EXCHANGE=$ex CTR_TRACE=gpurun_out/r02d_trace_n2_$ex.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_n2_$ex.json 2> gpurun_out/r02d_bench_n2_$ex.err
  echo "bench exit $?"; grep -v "^W\|^\[W\|warn" gpurun_out/r02d_bench_n2_$ex.err | tail -12 | cut -c1-300; cut -c1-700 gpurun_out/r02d_bench_n2_$ex.json
done

#!/bin/bash
# round 2, call L (N GPUs): exchange parity, then bench --gpus N on the peer-memory exchange
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${NGPU:-2}
echo "== parity worker ($N ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/p2p_worker.py > gpurun_out/r02l_parity_n$N.log 2>&1
echo "worker exit $?"; grep -E "PARITY|Error|error|Traceback" gpurun_out/r02l_parity_n$N.log | cut -c1-400 | tail -6
run() {  # tag exchange dist
  echo "== bench --gpus $N ($1)"
  CTR_SHARD_EXCHANGE=$2 CTR_TRACE=gpurun_out/r02l_trace_n${N}_$1.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 200 --warmup 5 --dist $3 ${BENCH_EXTRA:-} > gpurun_out/r02l_bench_n${N}_$1.json 2> gpurun_out/r02l_bench_n${N}_$1.err
  echo "bench exit $?"; grep -E "Error|error|Traceback|overflow" gpurun_out/r02l_bench_n${N}_$1.err | tail -5 | cut -c1-300; grep "^{" gpurun_out/r02l_bench_n${N}_$1.json | cut -c1-300
}
run p2p p2p uniform
if [ -n "${FULL:-}" ]; then
  run p2p_zipf p2p zipf
  run nccl nccl uniform
fi

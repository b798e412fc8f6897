"""torchrun worker for test_gpu_sharded: the sharded DeepFM train step on every exchange path
(peer-memory kernels and NCCL all-to-alls) against the fp64 oracle + TF-Adam, via
bench.sharded_parity_check (the same check bench.py --gpus N runs before timing)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    ok = True
    for exchange in ("p2p", "nccl"):
        res = bench.sharded_parity_check(rank, world, dev, exchange, B=192, steps=3)
        ok = ok and res["ok"]
        if rank == 0:
            print("PARITY %s" % json.dumps(res), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0 and ok:
        print("SHARDED-EXCHANGE-PARITY-OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

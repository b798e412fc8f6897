"""torchrun worker for test_gpu_sharded: sharded DeepFM (NCCL) vs the unsharded fp64 oracle.
Every rank evaluates its own local batch; the oracle evaluates the concatenated global batch."""
import os
import sys

import numpy as np
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
    import make_golden as mg
    from oracle import criteo, models as om
    from recsys_b200 import _core
    from recsys_b200 import criteo_schema as cs
    from recsys_b200.deepfm import deepfm
    from recsys_b200.estimator import VariableStore

    spec = mg.small_spec()
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(32, 16), seed=3)
    B = 96
    feats_all, batch_all = mg.model_batch("deepfm", B * world, 7, spec)
    sl = slice(rank * B, (rank + 1) * B)
    feats = {k: torch.from_numpy(np.asarray(v)[sl]) for k, v in feats_all.items()}
    labels = batch_all["labels"][sl]
    hb = [spec.rows[spec.fields.index(k)] for k in criteo.CAT]
    lin, emb = cs.build_columns(16, linear="indicator_all", hash_buckets=hb)
    params = {"linear_feature_columns": lin, "embedding_feature_columns": emb, "embedding_size": 16,
              "learning_rate": 1e-3, "dropout": 0.0, "deep_layers": "32,16", "device": dev,
              "variable_store": VariableStore(), "shard_embedding": True, "shard_slack": 4.0,
              "shard_exchange": "nccl"}      # backward() alone, then the gathered gradient
    m = params["variable_store"].get("deepfm", lambda: _core.DeepFMModel(params))
    m.load_state(p64)
    sp = deepfm.model_fn(feats, labels, "train", params)
    m.emb.check_overflow()
    # oracle: BN uses per-replica batch statistics (as MirroredStrategy does), so evaluate the
    # local batch; the global loss is the mean of the replica losses
    local_batch = {"rows": batch_all["rows"][sl], "labels": labels}
    out64, g64 = om.loss_and_grads("deepfm", p64, local_batch)
    logits = m.last["logits"].detach().cpu().double()
    rel = float(((logits - out64["logits"]).abs() / (out64["logits"].abs() + 0.1)).max())
    assert rel <= 1e-4, rel
    m.backward(m.last["loss"])
    # table gradient: owners hold sum over ranks of (1/world) * local gradient
    ge = (g64["emb"] / world).float().to(dev)
    dist.all_reduce(ge)
    full = m.emb.full_grad()
    err = float((full - ge).abs().max())
    assert err <= 1e-3 * float(ge.abs().max()) + 1e-7, err
    gd = (g64["dnn.0.w"] / world).float().to(dev)
    dist.all_reduce(gd)
    m._sync_dense_grads()
    errd = float((m.dense_grads()["dnn.0.w"] - gd).abs().max())
    assert errd <= 1e-3 * float(gd.abs().max()) + 1e-7, errd
    m.dense.grad.zero_()
    m.apply_gradients()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("SHARDED-PARITY-OK world=%d logits rel %.2e table grad err %.2e" % (world, rel, err))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

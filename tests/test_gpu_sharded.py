"""GPU: the device halves of the row-sharded exchange against their documented contracts
(1 GPU), and - when the box has >= 2 GPUs - the whole sharded DeepFM step over NCCL against
the unsharded fp64 oracle (launched through torch.distributed.run)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,G,cap", [(1, 2, 4), (1000, 2, 600), (159744, 8, 29952), (5000, 3, 100)])
def test_shard_bucket_contract(cuda, n, G, cap):
    from recsys_b200.sharded import CudaShardOps
    rng = np.random.default_rng(n)
    rows = torch.from_numpy(rng.integers(0, 10 ** 9, size=n).astype(np.int32)).to(cuda)
    send, slot, counts = CudaShardOps().bucket(rows, G, cap)
    rows_c, send_c, slot_c, counts_c = rows.cpu().numpy(), send.cpu().numpy(), slot.cpu().numpy(), \
        counts.cpu().numpy()
    assert np.array_equal(counts_c, np.bincount(rows_c % G, minlength=G))
    ok = slot_c >= 0
    assert ok.sum() == np.minimum(counts_c, cap).sum()
    assert len(np.unique(slot_c[ok])) == ok.sum()                      # one slot per lookup
    assert np.array_equal(slot_c[ok] // cap, rows_c[ok] % G)            # in the owner's slab
    assert np.array_equal(send_c[slot_c[ok]], rows_c[ok] // G)          # carries the local index
    used = np.zeros(G * cap, bool)
    used[slot_c[ok]] = True
    assert (send_c[~used] == -1).all()                                   # padding
    for o in range(G):                                                   # slabs fill from the front
        k = min(int(counts_c[o]), cap)
        assert used[o * cap:o * cap + k].all()


@pytest.mark.parametrize("D", [8, 16, 32])
def test_gather_and_scatter_rows(cuda, D):
    from recsys_b200.sharded import CudaShardOps
    ops = CudaShardOps()
    g = torch.Generator().manual_seed(D)
    R, n = 5000, 20011
    table = torch.randn(R, D, generator=g).to(cuda)
    w1 = torch.randn(R, generator=g).to(cuda)
    ids = torch.randint(-1, R, (n,), generator=g).int()
    ids[:50] = 7                                                        # duplicates on one row
    idc = ids.to(cuda)
    vec, w1v = ops.gather(table, w1, idc)
    ok = (ids >= 0)
    want = table.cpu()[ids.clamp(min=0).long()] * ok[:, None]
    assert torch.equal(vec.cpu(), want)
    assert torch.equal(w1v.cpu(), w1.cpu()[ids.clamp(min=0).long()] * ok)
    gr = torch.randn(n, D, generator=g)
    gw = torch.randn(n, generator=g)
    dt = torch.zeros(R, D, device=cuda)
    dw = torch.zeros(R, device=cuda)
    ops.scatter_add(idc, gr.to(cuda), gw.to(cuda), dt, dw)
    ref = torch.zeros(R, D, dtype=torch.float64).index_add_(0, ids[ok].long(), gr[ok].double())
    refw = torch.zeros(R, dtype=torch.float64).index_add_(0, ids[ok].long(), gw[ok].double())
    assert torch.allclose(dt.cpu().double(), ref, atol=1e-4)
    assert torch.allclose(dw.cpu().double(), refw, atol=1e-4)


def test_sharded_deepfm_matches_oracle_over_nccl(cuda):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200),
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "SHARDED-PARITY-OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_sharded_train_steps_match_oracle_on_both_exchanges(cuda):
    """>= 2 GPUs: three train steps of the row-sharded DeepFM through the peer-memory exchange
    kernels (csrc/p2p.cu) and through the NCCL all-to-alls, each against the fp64 oracle +
    TF-Adam on the gathered table (bench.sharded_parity_check)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2); bench.py --gpus N runs the same check "
                    "before timing and reports parity_ok")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 200),
           os.path.join(ROOT, "tests", "p2p_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "SHARDED-EXCHANGE-PARITY-OK" in res.stdout, \
        res.stdout[-3000:] + res.stderr[-3000:]

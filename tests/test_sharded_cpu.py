"""CPU, gloo, world_size 2: the host-side logic of the row-sharded table (slab capacity,
bucketing contract, the three all-to-alls, slot <-> lookup mapping, gradient routing back to
the owners).  The five device primitives are replaced by torch test doubles that implement
the documented C-ABI contracts (include/ctr_b200.h); the CUDA kernels themselves are checked
against the same contracts in tests/test_gpu_sharded.py."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class TorchShardOps:
    """Test doubles for ctr_shard_bucket / ctr_gather_rows / ctr_scatter_add_rows and the
    slot-indexed ctr_embed_fwd / ctr_embed_bwd."""

    def bucket(self, rows_flat, G, capacity):
        n = rows_flat.numel()
        send = torch.full((G * capacity,), -1, dtype=torch.int32)
        slot = torch.full((n,), -1, dtype=torch.int32)
        counts = torch.zeros(G, dtype=torch.int32)
        for i, r in enumerate(rows_flat.tolist()):
            o = r % G
            pos = int(counts[o])
            counts[o] += 1
            if pos < capacity:
                send[o * capacity + pos] = r // G
                slot[i] = o * capacity + pos
        return send, slot, counts

    def gather(self, table, w1, ids):
        idx = ids.long().clamp(min=0)
        vec = table[idx] * (ids >= 0).float()[:, None]
        w1v = w1[idx] * (ids >= 0).float() if w1 is not None else None
        return vec, w1v

    def scatter_add(self, ids, g, gw1, dtable, dw1):
        ok = ids >= 0
        dtable.index_add_(0, ids[ok].long(), g[ok])
        if dw1 is not None:
            dw1.index_add_(0, ids[ok].long(), gw1[ok])

    def interact_fwd(self, vec, w1v, slot2d, D, w1_fields, want_fm, want_y1, cross_w, cross_b):
        B, F = slot2d.shape
        E3 = vec[slot2d.long()]
        S = E3.sum(1)
        y2 = 0.5 * (S * S - (E3 * E3).sum(1)).sum(1) if want_fm else None
        mask = torch.tensor([(w1_fields >> f) & 1 for f in range(F)], dtype=torch.float32)
        y1 = (w1v[slot2d.long()] * mask).sum(1) if want_y1 else None
        return E3.reshape(B, F * D), S, y1, y2, None

    def interact_bwd(self, slot2d, dE, E, vec, S, dy2, dy1, w1_fields, D, n_slots):
        B, F = slot2d.shape
        g = torch.zeros(B, F, D) if dE is None else dE.view(B, F, D).clone()
        if dy2 is not None:
            g = g + dy2[:, None, None] * (S[:, None, :] - E.view(B, F, D))
        gsend = torch.zeros(n_slots, D)
        gsend.index_add_(0, slot2d.reshape(-1).long(), g.reshape(-1, D))
        gw1 = None
        if dy1 is not None:
            mask = torch.tensor([(w1_fields >> f) & 1 for f in range(F)], dtype=torch.float32)
            gw1 = torch.zeros(n_slots)
            gw1.index_add_(0, slot2d.reshape(-1).long(), (dy1[:, None] * mask).reshape(-1))
        return gsend, gw1


class TorchPackedShardOps(TorchShardOps):
    """Doubles for the packed exchange (one [D+4] slab record = row | w1 | pad per lookup, and
    the same for the gradients): the 3-all-to-all path that CudaShardOps takes."""
    packed = True

    def gather_packed(self, table, w1, ids):
        D = table.shape[1]
        vec, w1v = self.gather(table, w1, ids)
        slab = torch.zeros(ids.numel(), D + 4)
        slab[:, :D] = vec
        if w1v is not None:
            slab[:, D] = w1v
        return slab

    def scatter_add_packed(self, ids, gslab, D, dtable, dw1):
        self.scatter_add(ids, gslab[:, :D].contiguous(), gslab[:, D].contiguous(), dtable, dw1)

    def interact_fwd_packed(self, slab, slot2d, D, w1_fields, want_fm, want_y1, cross_w, cross_b,
                            want_lo=False):
        E, S, y1, y2, xl = self.interact_fwd(slab[:, :D], slab[:, D], slot2d, D, w1_fields, want_fm,
                                             want_y1, cross_w, cross_b)
        return E, S, y1, y2, xl, None

    def interact_bwd_packed(self, slot2d, dE, E, slab, S, dy2, dy1, w1_fields, D, n_slots):
        gsend, gw1 = self.interact_bwd(slot2d, dE, E, slab[:, :D], S, dy2, dy1, w1_fields, D, n_slots)
        gslab = torch.zeros(n_slots, D + 4)
        gslab[:, :D] = gsend
        if gw1 is not None:
            gslab[:, D] = gw1
        return gslab


def _worker(rank, world, port, q, packed=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recsys_b200 import feature_column as fc
        from recsys_b200 import sharded
        D, nrows = 16, [7, 50, 1001, 13]
        cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("k%d" % i, n), D)
                for i, n in enumerate(nrows)]
        lay = fc.layout(cols)
        F, R = lay.F, lay.total_rows
        g = torch.Generator().manual_seed(0)
        full_table = torch.randn(R, D, generator=g)
        full_w1 = torch.randn(R, generator=g)
        emb = sharded.ShardedFieldEmbedding(lay, torch.device("cpu"), with_w1=True, w1_fields=0b1011,
                                            shard_ops=TorchPackedShardOps() if packed else TorchShardOps(),
                                            capacity=34 * len(nrows))
        emb.load(full_table, full_w1)
        assert emb.table.shape[0] == (R - rank + world - 1) // world
        B = 33 + rank                                   # ragged: ranks hold different batch sizes
        rng = np.random.default_rng(10 + rank)
        rows = torch.from_numpy(np.stack([rng.integers(0, n, size=B) + o
                                          for n, o in zip(lay.rows, lay.offsets[:-1])], 1)).int()
        # (equal-split all-to-all: the slab capacity is fixed and identical on every rank)
        E, y1, y2, _ = emb.lookup(rows)
        # reference: the unsharded computation on the full table
        t = full_table.double().requires_grad_(True)
        w = full_w1.double().requires_grad_(True)
        Eo = t[rows.long()]
        So = Eo.sum(1)
        y2o = 0.5 * (So * So - (Eo * Eo).sum(1)).sum(1)
        mask = torch.tensor([1, 1, 0, 1], dtype=torch.float64)
        y1o = (w[rows.long()] * mask).sum(1)
        assert torch.allclose(E.double(), Eo.reshape(B, -1).detach(), atol=1e-6)
        assert torch.allclose(y2.double(), y2o.detach(), atol=1e-4)
        assert torch.allclose(y1.double(), y1o.detach(), atol=1e-5)
        gen = torch.Generator().manual_seed(rank)
        dE, dy1, dy2 = torch.randn(B, F * D, generator=gen), torch.randn(B, generator=gen), \
            torch.randn(B, generator=gen)
        torch.autograd.backward([E, y1, y2], [dE, dy1, dy2])
        loss = (Eo.reshape(B, -1) * dE.double()).sum() + (y1o * dy1.double()).sum() + \
            (y2o * dy2.double()).sum()
        loss.backward()
        # the owners hold the SUM over ranks of every rank's gradient
        tg = t.grad.float()
        wg = w.grad.float()
        dist.all_reduce(tg)
        dist.all_reduce(wg)
        full = emb.full_grad()
        assert torch.allclose(full, tg, atol=1e-4), float((full - tg).abs().max())
        assert torch.allclose(emb.dw1, wg[rank::world], atol=1e-4)
        emb.check_overflow()
        assert int(emb.counts.sum()) == B * F
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_world2(packed=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300 + (300 if packed else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, packed)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_slab_capacity_and_overflow_flag():
    from recsys_b200 import sharded
    assert sharded.slab_capacity(4096 * 39, 8, 1.5) % 4 == 0
    assert sharded.slab_capacity(4096 * 39, 8, 1.5) >= 4096 * 39 / 8 * 1.5
    ops = TorchShardOps()
    rows = torch.zeros(10, dtype=torch.int32)          # every lookup goes to owner 0
    send, slot, counts = ops.bucket(rows, 2, 4)
    assert counts.tolist() == [10, 0] and (slot[4:] == -1).all() and (send[4:] == -1).all()


def test_sharded_packed_exchange_world2():
    """The packed-slab routing (3 all-to-alls per step: ids, row|w1 slab, gradient slab)."""
    test_sharded_exchange_world2(packed=True)

"""Row-sharded embedding table over the GPUs of one box (BASELINE config 5).

The reference only replicates its tables (tf.distribute.MirroredStrategy, fm/fm.py:184-194);
sharding is the north-star extension for a table that exceeds one GPU's HBM.  One process per
GPU; rows are owned round-robin (owner = row % G, local index = row // G); dense weights are
replicated and their gradients all-reduced (NCCL).  Per step and rank:

  forward   bucket lookups by owner (ctr_shard_bucket, fixed-capacity slabs, no host sync)
            -> all-to-all ids -> owner gather (ctr_gather_rows) -> all-to-all vectors back
            -> ctr_embed_fwd over the received slab (slots play the role of row ids)
  backward  ctr_embed_bwd into a per-slot gradient slab -> all-to-all back to the owners
            -> ctr_scatter_add_rows into the local gradient accumulator -> Adam on touched rows

The exchange moves (G-1)/G * (4 + 64 + 64) B per lookup over NVLink; nothing else crosses.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import feature_column as fc
from . import ops
from .ops import _call, _p, _stream


class CudaShardOps:
    """The device halves of the exchange (libctr_b200)."""

    def bucket(self, rows_flat, G, capacity):
        n = rows_flat.numel()
        dev = rows_flat.device
        send = torch.empty(G * capacity, dtype=torch.int32, device=dev)
        slot = torch.empty(n, dtype=torch.int32, device=dev)
        counts = torch.empty(G, dtype=torch.int32, device=dev)
        _call("ctr_shard_bucket", _p(rows_flat), n, G, capacity, _p(send), _p(slot), _p(counts),
              _stream())
        return send, slot, counts

    # One exchange slab per direction: a looked-up row and its first-order weight travel as one
    # [D+4]-float record (row | w1 | pad), and so do their gradients on the way back: 3 all-to-alls
    # per step (ids, rows, gradients) instead of 5.  The fused lookup / scatter kernels read and
    # write the slabs in place through their row-stride parameters.
    packed = True

    def gather(self, table, w1, ids):
        n, D = ids.numel(), table.shape[1]
        vec = torch.empty((n, D), dtype=torch.float32, device=ids.device)
        w1v = torch.empty(n, dtype=torch.float32, device=ids.device) if w1 is not None else None
        _call("ctr_gather_rows", _p(table), _p(w1), _p(ids), n, D, _p(vec), _p(w1v),
              table.stride(0), w1.stride(0) if w1 is not None else 0, 0, 0, _stream())
        return vec, w1v

    def scatter_add(self, ids, g, gw1, dtable, dw1):
        _call("ctr_scatter_add_rows", _p(ids), _p(g), _p(gw1) if dw1 is not None else None,
              ids.numel(), dtable.shape[1], _p(dtable), _p(dw1), 0, 0, dtable.stride(0),
              dw1.stride(0) if dw1 is not None else 0, _stream())

    def gather_packed(self, table, w1, ids):
        n, D = ids.numel(), table.shape[1]
        P = D + 4
        slab = torch.empty((n, P), dtype=torch.float32, device=ids.device)
        _call("ctr_gather_rows", _p(table), _p(w1), _p(ids), n, D, _p(slab),
              _p(slab) + 4 * D if w1 is not None else None, table.stride(0),
              w1.stride(0) if w1 is not None else 0, P, P, _stream())
        return slab

    def scatter_add_packed(self, ids, gslab, D, dtable, dw1):
        P = D + 4
        _call("ctr_scatter_add_rows", _p(ids), _p(gslab), _p(gslab) + 4 * D if dw1 is not None else None,
              ids.numel(), D, _p(dtable), _p(dw1), P, P, dtable.stride(0),
              dw1.stride(0) if dw1 is not None else 0, _stream())

    def interact_fwd_packed(self, slab, slot2d, D, w1_fields, want_fm, want_y1, cross_w, cross_b,
                            want_lo=False):
        B, F = slot2d.shape
        dev = slot2d.device
        P = D + 4
        E = torch.empty((B, F * D), dtype=torch.float32, device=dev)
        E_lo = torch.empty_like(E) if want_lo else None
        S = torch.empty((B, D), dtype=torch.float32, device=dev) if want_fm else None
        y2 = torch.empty(B, dtype=torch.float32, device=dev) if want_fm else None
        y1 = torch.empty(B, dtype=torch.float32, device=dev) if want_y1 else None
        cross = cross_w is not None
        xl = torch.empty((B, F * D), dtype=torch.float32, device=dev) if cross else None
        _call("ctr_embed_fwd", _p(slab), _p(slab) + 4 * D if want_y1 else None, _p(slot2d), B, F, D,
              w1_fields, _p(E), _p(S), _p(y1), _p(y2), _p(cross_w), _p(cross_b),
              cross_w.shape[0] if cross else 0, _p(xl), _p(E_lo), P, P, _stream())
        return E, S, y1, y2, xl, E_lo

    def interact_bwd_packed(self, slot2d, dE, E, slab, S, dy2, dy1, w1_fields, D, n_slots):
        B, F = slot2d.shape
        P = D + 4
        gslab = torch.zeros((n_slots, P), dtype=torch.float32, device=slot2d.device)
        offs = (C.c_int64 * (F + 1))(*[1000 * f for f in range(F + 1)])
        _call("ctr_embed_bwd", _p(slot2d), _p(dE), _p(E), _p(slab), _p(S), _p(dy2), _p(dy1),
              w1_fields, offs, B, F, D, _p(gslab), _p(gslab) + 4 * D if dy1 is not None else None,
              P, P, _stream())
        return gslab

    def interact_fwd(self, vec, w1v, slot2d, D, w1_fields, want_fm, want_y1, cross_w, cross_b):
        B, F = slot2d.shape
        dev = slot2d.device
        E = torch.empty((B, F * D), dtype=torch.float32, device=dev)
        S = torch.empty((B, D), dtype=torch.float32, device=dev) if want_fm else None
        y2 = torch.empty(B, dtype=torch.float32, device=dev) if want_fm else None
        y1 = torch.empty(B, dtype=torch.float32, device=dev) if want_y1 else None
        cross = cross_w is not None
        xl = torch.empty((B, F * D), dtype=torch.float32, device=dev) if cross else None
        _call("ctr_embed_fwd", _p(vec), _p(w1v), _p(slot2d), B, F, D, w1_fields, _p(E), _p(S), _p(y1),
              _p(y2), _p(cross_w), _p(cross_b), cross_w.shape[0] if cross else 0, _p(xl), None, 0, 0, _stream())
        return E, S, y1, y2, xl

    def interact_bwd(self, slot2d, dE, E, vec, S, dy2, dy1, w1_fields, D, n_slots):
        B, F = slot2d.shape
        dev = slot2d.device
        gsend = torch.zeros((n_slots, D), dtype=torch.float32, device=dev)
        gw1 = torch.zeros(n_slots, dtype=torch.float32, device=dev) if dy1 is not None else None
        # slots are unique per lookup: no field is "tiny" (fake offsets 1000 apart)
        offs = (C.c_int64 * (F + 1))(*[1000 * f for f in range(F + 1)])
        _call("ctr_embed_bwd", _p(slot2d), _p(dE), _p(E), _p(vec), _p(S), _p(dy2), _p(dy1),
              w1_fields, offs, B, F, D, _p(gsend), _p(gw1), 0, 0, _stream())
        return gsend, gw1


def _a2a(x: torch.Tensor, group) -> torch.Tensor:
    out = torch.empty_like(x)
    dist.all_to_all_single(out, x, group=group)
    return out


def slab_capacity(n_lookups: int, G: int, slack: float) -> int:
    return int(math.ceil(n_lookups / G * slack / 4.0)) * 4


class ShardedFieldEmbedding:
    """Same surface as ops.FieldEmbedding (lookup / adam_step / dtable), rows split over the
    ranks of ``group``.  ``table`` holds this rank's rows only."""

    def __init__(self, lay: fc.Layout, device, group=None, with_w1=True, w1_fields=0,
                 adam_mode="lazy", seed=0, slack=1.5, shard_ops=None, capacity=None):
        self.lay, self.device, self.group = lay, device, group
        self.G = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.D, self.F, self.R = lay.dimension, lay.F, lay.total_rows
        self.R_local = (self.R - self.rank + self.G - 1) // self.G
        self.slack = slack
        self.fixed_capacity = capacity      # per-(src,dst) slab size; must agree on all ranks
        self.ops = shard_ops or CudaShardOps()
        g = torch.Generator(device=device).manual_seed(seed * 1000 + self.rank)
        D, RL = self.D, self.R_local
        self.with_w1, self.w1_fields = with_w1, w1_fields
        self.adam_mode = adam_mode
        # row records (ops.FieldEmbedding): theta | m | v | g | theta1 m1 v1 g1 | claim per owned row
        self.record = (adam_mode == "lazy" and device.type == "cuda"
                       and os.environ.get("CTR_ROW_RECORDS", "1") != "0")
        if self.record:
            S = 4 * D + 8
            self.rec = torch.zeros(RL, S, dtype=torch.float32, device=device)
            self.table, self._m = self.rec[:, 0:D], self.rec[:, D:2 * D]
            self._v, self.dtable = self.rec[:, 2 * D:3 * D], self.rec[:, 3 * D:4 * D]
            self.w1, self._m1 = self.rec[:, 4 * D], self.rec[:, 4 * D + 1]
            self._v1, self.dw1 = self.rec[:, 4 * D + 2], self.rec[:, 4 * D + 3]
            self._claim = self.rec[:, 4 * D + 4].view(torch.int32)
            if not with_w1:
                self.w1 = self.dw1 = self._m1 = self._v1 = None
        else:
            self.table = torch.empty(RL, D, dtype=torch.float32, device=device)
            self.dtable = torch.zeros_like(self.table)
            self.w1 = torch.empty(RL, dtype=torch.float32, device=device) if with_w1 else None
            self.dw1 = torch.zeros_like(self.w1) if with_w1 else None
            self._m = self._v = self._m1 = self._v1 = self._claim = None
        std = D ** -0.5
        step = 1 << 22
        for r0 in range(0, RL, step):
            blk = torch.empty(min(step, RL - r0), D, dtype=torch.float32, device=device)
            torch.nn.init.trunc_normal_(blk, std=std, a=-2 * std, b=2 * std, generator=g)
            self.table[r0:r0 + blk.shape[0]].copy_(blk)
        if with_w1:
            lim = math.sqrt(6.0 / (self.R + 1))
            self.w1.copy_((torch.rand(RL, generator=g, device=device) * 2 - 1) * lim)
        self._claim1 = None
        self.last_E_lo = None
        self._tag = 0
        self._anchor = torch.zeros((), device=device, requires_grad=True)
        self.recv_ids = None
        self.counts = None
        self.capacity = 0
        self.status = None      # the model's sticky status word (int32[1]); overflow sets bit 2

    # state -----------------------------------------------------------------------
    def load(self, table=None, w1=None):
        """Load a FULL [R, D] table (tests): each rank keeps rows rank, rank+G, ..."""
        with torch.no_grad():
            if table is not None:
                self.table.copy_(table[self.rank::self.G].to(self.device, torch.float32))
            if w1 is not None and self.with_w1:
                self.w1.copy_(w1.reshape(-1)[self.rank::self.G].to(self.device, torch.float32))

    def full_grad(self):
        """All-gather the sharded gradient accumulator into [R, D] (tests)."""
        rl = (self.R + self.G - 1) // self.G
        mine = torch.zeros(rl, self.D, dtype=torch.float32, device=self.device)
        mine[:self.R_local] = self.dtable
        parts = [torch.empty_like(mine) for _ in range(self.G)]
        dist.all_gather(parts, mine, group=self.group)
        full = torch.zeros(self.R, self.D, dtype=torch.float32, device=self.device)
        for r in range(self.G):
            n = (self.R - r + self.G - 1) // self.G
            full[r::self.G] = parts[r][:n]
        return full

    def check_overflow(self):
        """Raises if a slab overflowed in the last lookup (synchronises)."""
        if self.counts is not None and int(self.counts.max()) > self.capacity:
            raise RuntimeError("sharded exchange slab overflow: %d lookups for one owner, capacity "
                               "%d; raise slack" % (int(self.counts.max()), self.capacity))

    # forward / backward -----------------------------------------------------------
    def lookup(self, rows, want_fm=True, want_y1=True, cross_w=None, cross_b=None, want_lo=False):
        self._want_lo = bool(want_lo) and getattr(self.ops, "packed", False)
        return _ShardedEmbedFn.apply(self._anchor, self, rows, want_fm, want_y1 and self.with_w1,
                                     cross_w, cross_b)

    def zero_grad(self):
        self.dtable.zero_()
        if self.with_w1:
            self.dw1.zero_()

    def adam_step(self, rows_unused, lr_t, st):
        if self._m is None:
            self._m, self._v = torch.zeros_like(self.table), torch.zeros_like(self.table)
            self._claim = torch.zeros(self.R_local, dtype=torch.int32, device=self.device)
            if self.with_w1:
                self._m1, self._v1 = torch.zeros_like(self.w1), torch.zeros_like(self.w1)
        if self.adam_mode == "exact_tf":
            _call("ctr_adam_dense", _p(self.table), _p(self._m), _p(self._v), _p(self.dtable),
                  self.table.numel(), lr_t, st.beta1, st.beta2, st.eps, 1, st.state_ptr, 0, _stream())
            if self.with_w1:
                _call("ctr_adam_dense", _p(self.w1), _p(self._m1), _p(self._v1), _p(self.dw1),
                      self.w1.numel(), lr_t, st.beta1, st.beta2, st.eps, 1, st.state_ptr, 0, _stream())
            return
        self._tag += 1
        ids = self.recv_ids
        w = self.with_w1
        _call("ctr_adam_rows", _p(ids), ids.numel(), self.D, _p(self.table), _p(self._m),
              _p(self._v), _p(self.dtable), _p(self.w1) if w else None, _p(self._m1) if w else None,
              _p(self._v1) if w else None, _p(self.dw1) if w else None, _p(self._claim), self._tag,
              lr_t, st.beta1, st.beta2, st.eps, st.state_ptr, self.table.stride(0),
              self.w1.stride(0) if w else 0, self._claim.stride(0), _stream())


class _ShardedEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, emb: ShardedFieldEmbedding, rows, want_fm, want_y1, cross_w, cross_b):
        ctx.set_materialize_grads(False)
        B, F, D, G = rows.shape[0], emb.F, emb.D, emb.G
        rows = rows.contiguous()
        cap = emb.fixed_capacity or slab_capacity(B * F, G, emb.slack)
        send_ids, slot, counts = emb.ops.bucket(rows.view(-1), G, cap)
        recv_ids = _a2a(send_ids, emb.group)                       # ids this rank owns
        slot2d = slot.view(B, F)
        packed = getattr(emb.ops, "packed", False)
        if packed:
            slab = emb.ops.gather_packed(emb.table, emb.w1 if want_y1 else None, recv_ids)
            vec_back = _a2a(slab, emb.group)                       # [G*cap, D+4]: row | w1 | pad
            E, S, y1, y2, xl, emb.last_E_lo = emb.ops.interact_fwd_packed(
                vec_back, slot2d, D, emb.w1_fields, want_fm, want_y1, cross_w, cross_b,
                getattr(emb, "_want_lo", False))
        else:
            vec, w1v = emb.ops.gather(emb.table, emb.w1 if want_y1 else None, recv_ids)
            vec_back = _a2a(vec, emb.group)                        # [G*cap, D]
            w1_back = _a2a(w1v, emb.group) if want_y1 else None
            E, S, y1, y2, xl = emb.ops.interact_fwd(vec_back, w1_back, slot2d, D, emb.w1_fields,
                                                    want_fm, want_y1, cross_w, cross_b)
        emb.recv_ids, emb.counts, emb.capacity = recv_ids, counts, cap
        if emb.status is not None:       # sticky: bit 2 of the model's status word (see _core.py)
            emb.status.bitwise_or_((counts.max() > cap).to(torch.int32).mul_(4))
        ctx.emb, ctx.slot2d, ctx.recv_ids, ctx.E, ctx.S, ctx.vec = emb, slot2d, recv_ids, E, S, vec_back
        ctx.cross_w, ctx.cross_b = cross_w, cross_b
        cross = cross_w is not None
        ctx.flags = (want_fm, want_y1, cross, G * cap)
        ctx.packed = packed
        z = E.new_zeros(())
        outs = [E, y1 if want_y1 else z, y2 if want_fm else z, xl if cross else z]
        nd = [o for o, f in zip(outs[1:], (want_y1, want_fm, cross)) if not f]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dE, dy1, dy2, dxl):
        emb = ctx.emb
        want_fm, want_y1, cross, n_slots = ctx.flags
        D = emb.D
        dcw = dcb = None
        if cross and dxl is not None:
            L, W = ctx.cross_w.shape
            dx0 = torch.empty_like(ctx.E)
            dcw, dcb = torch.zeros_like(ctx.cross_w), torch.zeros_like(ctx.cross_b)
            _call("ctr_dcn_cross_bwd", _p(ctx.E), _p(ctx.cross_w), _p(ctx.cross_b), L, ctx.E.shape[0],
                  W, _p(dxl.contiguous()), _p(dx0), _p(dcw), _p(dcb), _stream())
            dE = dx0 if dE is None else dE + dx0
        dE = None if dE is None else dE.contiguous()
        dy2 = dy2.contiguous() if (want_fm and dy2 is not None) else None
        dy1 = dy1.contiguous() if (want_y1 and dy1 is not None) else None
        if dE is None and dy2 is None and dy1 is None:
            return None, None, None, None, None, dcw, dcb
        if dE is None and dy2 is None:
            dE = torch.zeros_like(ctx.E)
        if ctx.packed:
            gslab = emb.ops.interact_bwd_packed(ctx.slot2d, dE, ctx.E, ctx.vec, ctx.S, dy2, dy1,
                                                emb.w1_fields, D, n_slots)
            grecv = _a2a(gslab, emb.group)
            emb.ops.scatter_add_packed(ctx.recv_ids, grecv, D, emb.dtable,
                                       emb.dw1 if dy1 is not None else None)
            return None, None, None, None, None, dcw, dcb
        gsend, gw1 = emb.ops.interact_bwd(ctx.slot2d, dE, ctx.E, ctx.vec, ctx.S, dy2, dy1,
                                          emb.w1_fields, D, n_slots)
        grecv = _a2a(gsend, emb.group)
        gw1recv = _a2a(gw1, emb.group) if gw1 is not None else None
        emb.ops.scatter_add(ctx.recv_ids, grecv, gw1recv, emb.dtable,
                            emb.dw1 if gw1 is not None else None)
        return None, None, None, None, None, dcw, dcb


# ------------------------------------------------------------------- bench (N > 1)
def sharded_columns(total_rows: int, embedding_size: int, n_fields: int = 39):
    """Synthetic config 5: ``n_fields`` hashed fields sharing ``total_rows`` rows."""
    per = total_rows // n_fields
    lin, emb = [], []
    for i in range(n_fields):
        c = fc.categorical_column_with_hash_bucket("_c%d" % (i + 1), per)
        lin.append(fc.indicator_column(c))
        emb.append(fc.embedding_column(c, embedding_size))
    return lin, emb


def bench_config(args, world: int) -> dict:
    """The ``config`` object of the ``bench.py --gpus N`` line (N > 1), derived from the arguments
    and the environment only, so that the reference arm (``--impl reference``, rank 0 on the host
    cores) reports the very same workload."""
    B = args.batch
    total_rows = int(os.environ.get("CTR_SHARDED_ROWS", str(125_000_000 * world)))
    exchange = os.environ.get("CTR_SHARD_EXCHANGE", "p2p")
    slack = float(os.environ.get(
        "CTR_SHARD_SLACK", "1.1" if getattr(args, "dist", "uniform") == "uniform" else str(world)))
    return {"workload": "deepfm 39-field emb16, %d-row table row-sharded (row %% G) over %d "
                        "GPUs, %s of ids/vectors/grads, local batch %d, "
                        "fwd+bwd+Adam(lazy rows)" % (
                            total_rows, world,
                            "peer-memory exchange" if exchange == "p2p" else "NCCL all-to-all", B),
            "fields": 39, "embedding_size": 16, "batch_per_gpu": B, "global_batch": B * world,
            "table_rows": total_rows, "id_dist": getattr(args, "dist", "uniform"),
            "exchange": ("device-initiated over NVLink peer memory (cudaIpc arenas, flag "
                         "words; no NCCL call in the step; worst-case slabs)"
                         if exchange == "p2p" else
                         "3 NCCL all-to-alls per step (ids; row|w1 slab; gradient slab), "
                         "slab slack %.2f" % slack),
            "l2": "%.1f GB table shard per GPU > L2; distinct id batch every step"
                  % (total_rows / world * 64 / 1e9),
            "parallelism": "row-sharded table x%d + replicated dense weights (all-reduce)" % world}


def bench_main(args, rank, local, world):
    """DeepFM, 39 fields, emb 16, 1e9-row table row-sharded over ``world`` GPUs, local batch
    ``args.batch`` per GPU (weak scaling).  Rank 0 prints the JSON line."""
    import json
    import sys
    import time

    import numpy as np

    from .deepfm import deepfm
    from .estimator import VariableStore
    from .ops import PackedFeatures

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    # weak scaling: 125 M rows (36 GB of row records) per GPU, i.e. the 1 B-row table of BASELINE
    # config 5 at 8 GPUs; CTR_SHARDED_ROWS overrides the total
    total_rows = int(os.environ.get("CTR_SHARDED_ROWS", str(125_000_000 * world)))
    lin, emb = sharded_columns(total_rows, 16)
    params = {"linear_feature_columns": lin, "embedding_feature_columns": emb, "embedding_size": 16,
              "learning_rate": 1e-3, "dropout": 0.5, "deep_layers": "100,100", "device": dev,
              "variable_store": VariableStore(), "embedding_adam": "lazy", "shard_embedding": True,
              # exchange slabs sized 1.1x the mean per-(src,dst) load: uniform ids spread by < 1 %
              # (7 sigma = 0.05); check_overflow() below verifies that nothing was dropped
              # (Zipf ids put most lookups of a field on one owner: worst-case slabs there; the
              # peer-memory exchange always has worst-case slabs and ignores this)
              "shard_slack": float(os.environ.get(
                  "CTR_SHARD_SLACK", "1.1" if getattr(args, "dist", "uniform") == "uniform" else str(world))),
              "shard_exchange": os.environ.get("CTR_SHARD_EXCHANGE", "p2p"),
              "seed": 0}
    exchange = params["shard_exchange"]
    # sharded-vs-oracle parity on the live process group before anything is timed; the checker
    # lives with the bench (the package never imports the oracle)
    parity_fn = getattr(args, "parity_check", None)
    parity = parity_fn(rank, world, dev, exchange) if parity_fn is not None else \
        {"ok": None, "skipped": "no checker supplied"}
    lay = fc.layout(emb)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    rng = np.random.default_rng(1234 + rank)
    host, devb = [], []
    keys = [c.key for c in lay.columns]
    for _ in range(min(args.n_batches, 16)):
        if getattr(args, "dist", "uniform") == "zipf":     # hot rows: Zipf(1.05) clipped to the field
            cols = [np.minimum(rng.zipf(1.05, size=B) - 1, n - 1) for n in lay.rows]
        else:
            cols = [rng.integers(0, n, size=B) for n in lay.rows]
        cat = torch.from_numpy(np.stack(cols, 1).astype(np.int64)).pin_memory()
        cont = torch.zeros((B, 0), dtype=torch.float32).pin_memory()
        lab = torch.from_numpy((rng.random((B, 1)) < 0.22).astype(np.float32)).pin_memory()
        host.append((PackedFeatures(cont, cat, [], keys), lab))
        devb.append((PackedFeatures(cont.to(dev), cat.to(dev), [], keys), lab.to(dev)))
    from .estimator import GraphedTrainStep
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))   # one (high-priority) stream
    sp = deepfm.model_fn(devb[0][0], devb[0][1], "train", params)      # creates the variables
    sp.train_op()
    n0 = ops.LAUNCHES["n"]
    sp = deepfm.model_fn(devb[0][0], devb[0][1], "train", params)
    sp.train_op()
    per_step_launches = ops.LAUNCHES["n"] - n0
    del sp
    model = params["variable_store"]._objs["deepfm"]
    torch.cuda.synchronize()
    dist.barrier()
    graphed, mode = None, "eager"
    if not args.eager:
        try:     # NCCL collectives are captured with the kernels; every rank captures the same graph
            graphed = GraphedTrainStep(deepfm.model_fn, params, host[0][0], host[0][1], warmup=3)
            mode = "whole step incl. NCCL all-to-all captured in one CUDA graph per rank"
        except Exception as e:
            sys.stderr.write("rank %d: graph capture failed, eager: %r\n" % (rank, e))
            mode = "eager (graph capture failed: %s)" % str(e)[:100]
    ok = torch.tensor([1 if graphed is not None else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok) == 0:
        graphed = None

    def step(batch):
        f, l = batch
        if graphed is not None:
            return graphed(f, l)
        s = deepfm.model_fn(f, l, "train", params)
        s.train_op()
        return s.loss

    def timed(batches, read_loss):
        for i in range(W):
            step(batches[i % len(batches)])
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        slot = torch.zeros(K, dtype=torch.float32).pin_memory()
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(K):
            loss = step(batches[(W + i) % len(batches)])
            if read_loss:
                slot[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1), wall], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    clk = None
    try:                                            # clocks / throttle reasons during the timed region
        cs = getattr(args, "clock_sampler", None)
        clk = cs(local).__enter__() if (cs is not None and rank == 0) else None
    except Exception:
        clk = None
    ms, _ = timed(devb, False)
    ms_e2e, wall_e2e = timed(host, True)
    clocks = None
    if clk is not None:
        try:
            clk.__exit__(None, None, None)
            clocks = clk.summary()
        except Exception:
            clocks = None
    model.emb.check_overflow()
    model.check_status()
    if os.environ.get("CTR_TRACE"):         # in-graph kernel timeline of a few steps (every rank
        from torch.profiler import ProfilerActivity, profile   # runs them; rank 0 prints)
        dist.barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(6):          # device-resident batches: least skew between the ranks
                step(devb[i % len(devb)])
            torch.cuda.synchronize()
        if rank == 0:
            agg = {}
            for e in prof.events():
                if e.device_type == torch.autograd.DeviceType.CUDA:
                    a = agg.setdefault(e.name[:64], [0.0, 0])
                    a[0] += e.device_time
                    a[1] += 1
            lines = ["%8.2f us x %4.1f/step  %s" % (t / n, n / 6, k)
                     for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])]
            for ln in lines:
                sys.stderr.write("TRACE " + ln + "\n")
            out = os.environ.get("CTR_TRACE")
            if out not in ("1", ""):
                with open(out, "w") as fh:
                    fh.write("in-graph kernel times of the sharded DeepFM step, rank 0 of %d, %s exchange, "
                             "%.3f ms/step (CUPTI over 6 graph replays; a kernel that waits on a peer's "
                             "flag - p2p_gather_reply, p2p_scatter, p2p_adam_dense, the lookup - includes "
                             "the wait, and the first replay absorbs the ranks' start-up skew)\n"
                             % (world, exchange, ms / K) + "\n".join(lines) + "\n")
        dist.barrier()
    if rank == 0:
        f, l = host[0]
        G = world
        line = {
            "metric": "CTR samples/sec (Criteo 39-field emb16)", "value": K * B * world / (ms / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": bench_config(args, world),
            "e2e": {"value": K * B * world / (max(ms_e2e, wall_e2e) / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": (f.cat.numel() * 8 + l.numel() * 4) * world,
                    "d2h_bytes_per_step": 4 * world,
                    "api": "estimator.GraphedTrainStep(deepfm.model_fn, params)(pinned PackedFeatures, "
                           "labels)" if graphed is not None else
                           "deepfm.model_fn(pinned PackedFeatures, labels, 'train', params).train_op()"},
            "clocks": clocks,
            "gpu_launches": per_step_launches * K * world, "gpu_launches_per_step": per_step_launches,
            "launch_mode": mode,
            "parity_ok": parity["ok"], "parity": parity,
        }
        # NVLink roofline: bytes this GPU must send per step (ids + its share of rows it owns +
        # gradients, each (G-1)/G remote; P = 20 floats per record; + the dense gradient to G-1
        # peers) over the measured 770 GB/s per direction (B200_PROFILING.md)
        rec_b = 80
        n_dense = int(model.dense.numel)
        nv = (G - 1) / G * B * 39 * (4 + rec_b + rec_b) + (G - 1) * n_dense * 4
        line["nvlink_bytes_per_gpu_per_step"] = int(nv)
        line["roofline"] = {"bound": "nvlink", "achieved": nv / (ms / K * 1e-3) / 1e9, "peak": 770.0,
                            "unit": "GB/s", "frac": nv / (ms / K * 1e-3) / 1e9 / 770.0, "traffic": None,
                            "peak_source": "measured peer copy, one direction per GPU "
                                           "(B200_PROFILING.md; 900 nominal)",
                            "kernel": "p2p_gather_reply + p2p_grad_send + p2p_dense_push (stores "
                                      "into peer HBM)" if exchange == "p2p" else "NCCL all_to_all",
                            "note": "the step is latency bound (small messages): the fraction says how "
                                    "far the wire is from being the limit",
                            "whole_step_hbm": {"achieved": 8276.0 * B / (ms / K * 1e-3) / 1e9,
                                               "unit": "GB/s per GPU (algorithmic 8276 B/sample)"}}
        if not getattr(args, "no_cpu_baseline", False) and getattr(args, "time_oracle", None):
            try:
                n_cpu, el, cores = args.time_oracle("deepfm", B, "ref", "uniform",
                                                    max_seconds=min(args.cpu_seconds, 10.0))
                line["cpu_baseline"] = {
                    "value": n_cpu * B / el, "unit": "samples/s", "cores": cores, "kind": "port",
                    "sample": "%d fwd+bwd steps of batch %d (%.1f s) of the oracle's torch-CPU fp32 "
                              "deepfm on the reference-capped table (the 1e9-row table does not fit "
                              "host memory twice); rank 0 only; no optimizer step" % (n_cpu, B, el)}
            except Exception as e:        # the baseline is informative, never fatal
                line["cpu_baseline"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    # Tearing NCCL down under live CUDA graphs that captured its collectives can hang at exit:
    # agree that everyone is done, then leave without running the destructors.
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)

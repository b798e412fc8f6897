"""Row-sharded embedding table with a device-initiated exchange over NVLink peer memory
(csrc/p2p.cu; BASELINE config 5).

One process per GPU.  ``P2PArena`` allocates this rank's arena through the C ABI
(ctr_p2p_alloc), trades the cudaIpc handles over the process group and maps every peer's arena;
``P2PShardedEmbedding`` has the surface of ``ops.FieldEmbedding`` / ``sharded.ShardedFieldEmbedding``
and runs a step as the kernel chain K1..K6 of include/ctr_b200.h with no NCCL call inside:
ids, looked-up rows and gradients are stored straight into the consumer's HBM, flags carry the
step number, the owner applies scatter-add and Adam in one pass, and the replicated dense weights
are reduced by every rank summing the G gradient copies its peers stored into its arena.
torch.distributed is used once, at construction, to trade the handles.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from . import feature_column as fc
from .ops import _call, _p, _stream


def _align(n, a=256):
    return (n + a - 1) // a * a


class P2PArena:
    """This rank's exchange arena + the mapped arenas of its peers (ctr_p2p_ctx)."""

    def __init__(self, device, capacity: int, D: int, n_dense: int, group=None,
                 spin_limit_ms: int = 10000):
        lib = _lib.load()
        self.device, self.group = device, group
        G, me = dist.get_world_size(group), dist.get_rank(group)
        if G > 8:
            raise ValueError("the peer-memory exchange is built for one NVSwitch box (<= 8 ranks)")
        P = D + 4
        self.G, self.me, self.capacity, self.P = G, me, int(capacity), P
        self.n_dense = _align(max(int(n_dense), 4), 4)
        off, o = {}, 256                                   # [0, 256): header {step, err}
        for name, n_int in (("req_flag", 2 * G), ("req_cnt", 2 * G), ("resp_flag", G),
                            ("grad_flag", G), ("dense_flag", G), ("counts", G), ("sent", G),
                            ("done", 1 + G)):
            off[name] = o
            o = _align(o + 4 * n_int)
        for name, nbytes in (("req_ids", 2 * G * self.capacity * 4),
                             ("inv", G * self.capacity * 4),
                             ("resp", G * self.capacity * P * 4),
                             ("grad", G * self.capacity * P * 4),
                             ("dense", G * self.n_dense * 4)):
            off[name] = o
            o = _align(o + nbytes)
        self.off, self.nbytes = off, o
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.check(lib.ctr_p2p_alloc(o, C.byref(ptr), C.addressof(handle)))
        self.ptr = ptr.value
        # trade the handles (the one collective of this path)
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
        parts = [torch.empty_like(mine) for _ in range(G)]
        dist.all_gather(parts, mine, group=group)
        self.peers = []
        for r in range(G):
            if r == me:
                self.peers.append(self.ptr)
                continue
            h = (C.c_ubyte * 64)(*parts[r].cpu().tolist())
            pp = C.c_void_p()
            _lib.check(lib.ctr_p2p_open(C.addressof(h), C.byref(pp)))
            self.peers.append(pp.value)
        c = _lib.P2PCtx()
        for r in range(8):
            c.peer[r] = self.peers[r] if r < G else None
        c.me, c.G, c.capacity, c.record_floats = me, G, self.capacity, P
        c.off_req_flag, c.off_req_cnt = off["req_flag"], off["req_cnt"]
        c.off_resp_flag, c.off_grad_flag, c.off_dense_flag = off["resp_flag"], off["grad_flag"], off["dense_flag"]
        c.off_req_ids, c.off_resp, c.off_grad, c.off_dense = off["req_ids"], off["resp"], off["grad"], off["dense"]
        c.off_counts, c.off_done = off["counts"], off["done"]
        c.off_inv, c.off_sent = off["inv"], off["sent"]
        c.n_dense, c.spin_limit_ms = self.n_dense, int(spin_limit_ms)
        self.ctx = c
        dist.barrier(group=group)          # every arena is mapped before anyone stores into one

    @property
    def ref(self):
        return C.byref(self.ctx)

    def status(self):
        """(step, err): synchronises.  err bit 0 = a bounded wait timed out (a peer never came)."""
        step, err = C.c_int32(), C.c_int32()
        _lib.check(_lib.load().ctr_p2p_status(self.ref, C.addressof(step), C.addressof(err)))
        return step.value, err.value

    def check(self):
        step, err = self.status()
        if err:
            raise RuntimeError("peer-memory exchange: a wait timed out at step %d (a rank did not "
                               "run the same step)" % step)

    def close(self):
        lib = _lib.load()
        for r, p in enumerate(self.peers):
            if r != self.me and p:
                lib.ctr_p2p_close(p)
        if self.ptr:
            lib.ctr_p2p_free(self.ptr)
        self.peers, self.ptr = [], None


class P2PShardedEmbedding:
    """Same surface as ops.FieldEmbedding, rows split over the ranks (owner = row % G, local index
    = row // G), exchanged over peer memory.  The optimiser of the table runs INSIDE the backward
    (scatter-add + Adam in one pass on the owner), so a train step is
    ``arm_fused(...)`` -> backward; ``adam_step`` then only acknowledges it."""

    p2p = True

    def __init__(self, lay: fc.Layout, device, group=None, with_w1=True, w1_fields=0, seed=0,
                 spin_limit_ms: int = 10000):
        self.lay, self.device, self.group = lay, device, group
        self.G, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.D, self.F, self.R = lay.dimension, lay.F, lay.total_rows
        self.R_local = (self.R - self.rank + self.G - 1) // self.G
        D, RL = self.D, self.R_local
        self.with_w1, self.w1_fields = with_w1, w1_fields
        self.adam_mode = "lazy"
        self.record = True
        S = 4 * D + 8
        self.rec = torch.zeros(RL, S, dtype=torch.float32, device=device)
        self.table, self._m = self.rec[:, 0:D], self.rec[:, D:2 * D]
        self._v, self.dtable = self.rec[:, 2 * D:3 * D], self.rec[:, 3 * D:4 * D]
        self.w1, self.dw1 = self.rec[:, 4 * D], self.rec[:, 4 * D + 3]
        self.ld = S
        g = torch.Generator(device=device).manual_seed(seed * 1000 + self.rank)
        std = D ** -0.5
        step = 1 << 22
        for r0 in range(0, RL, step):
            blk = torch.empty(min(step, RL - r0), D, dtype=torch.float32, device=device)
            torch.nn.init.trunc_normal_(blk, std=std, a=-2 * std, b=2 * std, generator=g)
            self.table[r0:r0 + blk.shape[0]].copy_(blk)
        if with_w1:
            lim = math.sqrt(6.0 / (self.R + 1))
            self.w1.copy_((torch.rand(RL, generator=g, device=device) * 2 - 1) * lim)
        # the arena is sized by the first batch (lookups per rank) and the model's dense
        # parameter count (``n_dense``, set by the model before the first step); creating it is
        # a collective, and so is growing it
        self.arena: Optional[P2PArena] = None
        self.n_dense = 4
        self.spin_limit_ms = spin_limit_ms
        self.last_E_lo = None
        self._anchor = torch.zeros((), device=device, requires_grad=True)
        self._fused = None
        self._fused_done = False
        self.can_fuse = True
        self._want_lo = False
        self._count = False
        self._side = None
        self._pre_scatter = None      # main-stream event: every dense gradient is complete

    # state -----------------------------------------------------------------------
    def load(self, table=None, w1=None):
        """Load a FULL [R, D] table (tests): each rank keeps rows rank, rank+G, ..."""
        with torch.no_grad():
            if table is not None:
                self.table.copy_(table[self.rank::self.G].to(self.device, torch.float32))
            if w1 is not None and self.with_w1:
                self.w1.copy_(w1.reshape(-1)[self.rank::self.G].to(self.device, torch.float32))

    def _full(self, local, width):
        rl = (self.R + self.G - 1) // self.G
        mine = torch.zeros(rl, width, dtype=torch.float32, device=self.device)
        mine[:self.R_local] = local.reshape(self.R_local, width)
        parts = [torch.empty_like(mine) for _ in range(self.G)]
        dist.all_gather(parts, mine, group=self.group)
        full = torch.zeros(self.R, width, dtype=torch.float32, device=self.device)
        for r in range(self.G):
            n = (self.R - r + self.G - 1) // self.G
            full[r::self.G] = parts[r][:n]
        return full

    def full_table(self):
        """All-gather the shards into [R, D] (tests; a collective)."""
        return self._full(self.table, self.D)

    def full_w1(self):
        return self._full(self.w1, 1).reshape(-1)

    def ensure_arena(self, lookups: int) -> P2PArena:
        if self.arena is None or self.arena.capacity < lookups or self.arena.n_dense < self.n_dense:
            if self.arena is not None:
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)
                self.arena.close()
                # the new arena counts its steps from zero again: forget the old step tags
                self.rec[:, 4 * self.D + 4].zero_()
            self.arena = P2PArena(self.device, lookups, self.D, self.n_dense, self.group,
                                  self.spin_limit_ms)
        return self.arena

    def check_overflow(self):
        """Slabs are sized for the worst case; what can go wrong is a peer that never arrives."""
        if self.arena is not None:
            self.arena.check()

    # forward / backward -----------------------------------------------------------
    def lookup(self, rows, want_fm=True, want_y1=True, cross_w=None, cross_b=None, want_lo=False,
               training=False):
        self._want_lo = bool(want_lo)
        self._count = bool(training)
        return _P2PEmbedFn.apply(self._anchor, self, rows, want_fm, want_y1 and self.with_w1,
                                 cross_w, cross_b)

    def arm_fused(self, rows, lr_t, st) -> bool:
        self._fused = (lr_t, st)
        return True

    def adam_step(self, rows_unused, lr_t, st):
        if self._fused_done:
            self._fused_done = False
            return
        raise RuntimeError("the peer-memory sharded table applies its optimiser inside the "
                           "backward of a train step (spec.train_op()); a bare backward() + "
                           "apply_gradients() is not supported on this path")

    # replicated dense weights ------------------------------------------------------
    def dense_step(self, dense, lr_t, st, tower=None):
        """All-reduce + Adam of the replicated dense parameters without NCCL: push this rank's
        gradient into every arena, then Adam over the sum of the G copies.  It runs on a side
        stream, BESIDE the gradient exchange and the owner-side optimiser of the table (K4, K5):
        it only depends on the dense gradients, which are complete before K4 starts (event
        ``_pre_scatter``) and once the tower's weight-gradient kernels are done.  The Adam schedule
        is advanced on the main stream after both have finished (K5 reads lr_t from it)."""
        a = self.arena
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        side = tower.side if tower is not None else self._side
        if self._pre_scatter is not None:
            side.wait_event(self._pre_scatter)
            self._pre_scatter = None
        else:
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        if tower is not None and tower._pending is not None:
            side.wait_event(tower._pending)
            tower._pending = None
        with torch.cuda.stream(side):
            _call("ctr_p2p_dense_push", _p(dense.grad), dense.numel, a.ref, side.cuda_stream)
            _call("ctr_p2p_adam_dense", _p(dense.flat), _p(dense.m), _p(dense.v), _p(dense.grad),
                  dense.numel, lr_t, st.beta1, st.beta2, st.eps, st.state_ptr, 0, a.ref,
                  side.cuda_stream)
            done = torch.cuda.Event()
            done.record(side)
        main.wait_event(done)
        st.advance()                      # t += 1, lr_t of the next step (ctr_adam_tick)


class _P2PEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, emb: P2PShardedEmbedding, rows, want_fm, want_y1, cross_w, cross_b):
        ctx.set_materialize_grads(False)
        B, F, D = rows.shape[0], emb.F, emb.D
        a = emb.ensure_arena(B * F)
        dev = rows.device
        rows = rows.contiguous()
        slot = torch.empty((B, F), dtype=torch.int32, device=dev)
        _call("ctr_p2p_bucket_send", _p(rows), B * F, a.ref, _p(slot), _stream())             # K1
        _call("ctr_p2p_gather_reply", _p(emb.rec), emb.ld, D, 1 if want_y1 else 0,
              1 if emb._count else 0, a.ref, _stream())                                       # K2
        E = torch.empty((B, F * D), dtype=torch.float32, device=dev)
        E_lo = torch.empty_like(E) if emb._want_lo else None
        S = torch.empty((B, D), dtype=torch.float32, device=dev) if want_fm else None
        y2 = torch.empty(B, dtype=torch.float32, device=dev) if want_fm else None
        y1 = torch.empty(B, dtype=torch.float32, device=dev) if want_y1 else None
        cross = cross_w is not None
        xl = torch.empty((B, F * D), dtype=torch.float32, device=dev) if cross else None
        _call("ctr_embed_fwd_p2p", _p(slot), B, F, D, emb.w1_fields, 1 if want_y1 else 0, _p(E),
              _p(S), _p(y1), _p(y2), _p(cross_w), _p(cross_b), cross_w.shape[0] if cross else 0,
              _p(xl), _p(E_lo), a.ref, _stream())                                             # K3
        emb.last_E_lo = E_lo
        ctx.emb, ctx.slot, ctx.E, ctx.S = emb, slot, E, S
        ctx.cross_w, ctx.cross_b = cross_w, cross_b
        ctx.flags = (want_fm, want_y1, cross)
        z = E.new_empty(())
        outs = [E, y1 if want_y1 else z, y2 if want_fm else z, xl if cross else z]
        nd = [o for o, f in zip(outs[1:], (want_y1, want_fm, cross)) if not f]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dE, dy1, dy2, dxl):
        emb = ctx.emb
        want_fm, want_y1, cross = ctx.flags
        B, F, D = ctx.slot.shape[0], emb.F, emb.D
        a = emb.arena
        dcw = dcb = None
        if cross and dxl is not None:
            L, W = ctx.cross_w.shape
            dx0 = torch.empty_like(ctx.E)
            dcw, dcb = torch.zeros_like(ctx.cross_w), torch.zeros_like(ctx.cross_b)
            _call("ctr_dcn_cross_bwd", _p(ctx.E), _p(ctx.cross_w), _p(ctx.cross_b), L, B, W,
                  _p(dxl.contiguous()), _p(dx0), _p(dcw), _p(dcb), _stream())
            dE = dx0 if dE is None else dE + dx0
        dE = None if dE is None else dE.contiguous()
        dy2 = dy2.contiguous() if (want_fm and dy2 is not None) else None
        dy1 = dy1.contiguous() if (want_y1 and dy1 is not None) else None
        fused, emb._fused = emb._fused, None
        if fused is None:
            raise RuntimeError("the peer-memory sharded table needs spec.train_op() (its optimiser "
                               "runs inside the backward); a bare backward() is not supported")
        if dE is None and dy2 is None:
            dE = torch.zeros_like(ctx.E)
        lr_t, st = fused
        emb._pre_scatter = torch.cuda.Event()
        emb._pre_scatter.record(torch.cuda.current_stream())
        _call("ctr_p2p_grad_send", _p(ctx.slot), _p(dE), _p(ctx.S) if dy2 is not None else None,
              _p(dy2), _p(dy1), emb.w1_fields, B, F, D, a.ref, _stream())                     # K4
        _call("ctr_p2p_scatter_adam", _p(emb.rec), emb.ld, D, 1 if emb.with_w1 else 0,
              1 if want_fm else 0, lr_t, st.beta1, st.beta2, st.eps, st.state_ptr, a.ref,
              _stream())                                                                      # K5
        emb._fused_done = True
        return None, None, None, None, None, dcw, dcb

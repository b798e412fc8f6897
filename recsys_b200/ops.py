"""Host-side operators over libctr_b200: id pipeline, fused multi-field lookup
(+FM / +DCN cross) with its scatter-add backward, TF-Adam, and the flat
dense-parameter store.  PyTorch supplies device memory, streams and autograd
glue only; every op here launches hand-written sm_100a kernels through the C
ABI and raises if the library is missing or the device is not a B200.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from . import feature_column as fc

LAUNCHES = {"n": 0}   # kernels launched through the C ABI (bench.py reports it)


def _call(name, *args):
    lib = _lib.load()
    LAUNCHES["n"] += 1
    _lib.check(getattr(lib, name)(*args))


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: recsys_b200 has no CPU path" % what)


# ------------------------------------------------------------------ id pipeline
class PackedFeatures(dict):
    """A features dict whose per-key tensors are column views of two packed
    buffers - ``cont`` f32 [B, n_cont] and ``cat`` i64 [B, n_cat] - so that the
    model_fn can move a batch with two copies instead of 39."""

    def __init__(self, cont: torch.Tensor, cat: torch.Tensor, cont_keys: Sequence[str],
                 cat_keys: Sequence[str]):
        super().__init__()
        self.cont, self.cat = cont, cat
        self.cont_keys, self.cat_keys = list(cont_keys), list(cat_keys)
        # optional: the [B, F] row ids of this batch, computed ahead of the step
        # (estimator.GraphedTrainStep runs the model's ``prefetch_ids`` on its copy stream)
        self.rows = None
        for j, k in enumerate(self.cont_keys):
            self[k] = cont[:, j:j + 1]
        for j, k in enumerate(self.cat_keys):
            self[k] = cat[:, j:j + 1]


class IdPipeline:
    """features -> rows[B,F] int32 on the device (ctr_criteo_rows).  Replaces the
    id half of ``input_layer`` (fm/fm.py:76-80,89) for log-bucketised numerics and
    pre-hashed categoricals; raw byte strings are hashed by ``hash_strings``."""

    def __init__(self, lay: fc.Layout, device):
        self.lay = lay
        self.device = device
        self.cont_keys: List[str] = []
        self.cat_keys: List[str] = []
        self.raw_int64: Dict[str, bool] = {}      # columns fed raw int64 keys (hashed here)
        descs = (_lib.FieldDesc * lay.F)()
        bnd: List[float] = []
        for f, col in enumerate(lay.columns):
            cc = col.categorical_column
            d = descs[f]
            d.n_rows, d.row_offset = lay.rows[f], lay.offsets[f]
            if isinstance(cc, fc.BucketizedColumn):
                d.kind, d.src = 0, len(self.cont_keys)
                d.bnd_begin, d.bnd_count = len(bnd), len(cc.boundaries)
                off = cc.source_column.log_offset
                if off is None:
                    raise ValueError("numeric column %s needs a log(x+c) normalizer" % cc.key)
                d.log_offset = off
                bnd.extend(cc.boundaries)
                self.cont_keys.append(cc.key)
            else:
                d.kind, d.src = 1, len(self.cat_keys)
                self.cat_keys.append(cc.key)
                self.raw_int64[cc.key] = getattr(cc, "dtype", "string") == "int64"
        raw = bytes(descs)
        self.fields_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.bnd_dev = torch.tensor(bnd if bnd else [0.0], dtype=torch.float32, device=device)
        self.n_bnd = len(bnd)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.n_buckets_dev = torch.tensor(lay.rows, dtype=torch.int32, device=device)
        self.row_offset_dev = torch.tensor(lay.offsets[:-1], dtype=torch.int32, device=device)

    def _raw_order(self):
        """Column permutations from the decoder's raw order (_c1.._c13 / _c14.._c39) to this
        pipeline's packed order, and the bucket counts in raw order (device tensors, cached)."""
        if getattr(self, "_raw_perm", None) is None:
            ci = [int(k[2:]) - 1 for k in self.cont_keys]
            ki = [int(k[2:]) - 14 for k in self.cat_keys]
            nb = [1] * 26
            for k, j in zip(self.cat_keys, ki):
                nb[j] = self.lay.rows[self.lay.field_of(k)]
            dev = self.device
            self._raw_perm = (torch.tensor(ci, dtype=torch.long, device=dev),
                         torch.tensor(ki, dtype=torch.long, device=dev),
                         torch.tensor(nb, dtype=torch.int32, device=dev))
        return self._raw_perm

    def pack_raw(self, batch) -> tuple:
        """A ``data.CriteoRawBatch`` (decoder output: numerics + fixed-width string slots) ->
        (cont, cat) on the device; the strings are fingerprinted there (ctr_hash_slots)."""
        ci, ki, nb = self._raw_order()
        dev = self.device
        cont = batch.cont.to(dev, non_blocking=True).index_select(1, ci) if self.cont_keys else None
        cat = None
        if self.cat_keys:
            cb = batch.cat_bytes.to(dev, non_blocking=True)
            cl = batch.cat_len.to(dev, non_blocking=True)
            B, nf, slot = cb.shape
            ids = torch.empty((B, nf), dtype=torch.int64, device=dev)
            _call("ctr_hash_slots", _p(cb), slot, _p(cl), B * nf, nf, _p(nb), _p(ids), _stream())
            cat = ids.index_select(1, ki)
        return cont, cat

    def pack(self, features) -> tuple:
        """-> (cont f32 [B,n_cont], cat i64 [B,n_cat]) on the device."""
        if hasattr(features, "cat_bytes") and hasattr(features, "cat_len"):
            return self.pack_raw(features)
        if isinstance(features, PackedFeatures) and features.cont_keys == self.cont_keys \
                and features.cat_keys == self.cat_keys and not any(self.raw_int64.values()):
            cont, cat = features.cont, features.cat
        else:
            cont = torch.cat([torch.as_tensor(features[k]).reshape(-1, 1).float()
                              for k in self.cont_keys], 1) if self.cont_keys else None
            cats = []
            for k in self.cat_keys:
                v = features[k]
                if not torch.is_tensor(v) and not _is_int_array(v):
                    v = self.hash_strings_host(k, v)
                elif self.raw_int64[k]:
                    # categorical_column_with_hash_bucket(dtype=int64): raw keys, hashed as their
                    # decimal strings on the device (deepfm/deepfm.py:41,46) [TF-sem]
                    v = hash_int64(torch.as_tensor(v), self.lay.rows[self.lay.field_of(k)],
                                   self.device)
                else:
                    v = torch.as_tensor(v)
                cats.append(v.reshape(-1, 1).long())
            cat = torch.cat(cats, 1) if cats else None
        if cont is not None:
            cont = cont.to(self.device, non_blocking=True).contiguous()
        if cat is not None:
            cat = cat.to(self.device, non_blocking=True).contiguous()
        return cont, cat

    def hash_strings_host(self, key, values) -> torch.Tensor:
        """Raw byte strings of one categorical column -> local ids (device hash)."""
        f = self.lay.field_of(key)
        flat = [bytes(v) for v in _flatten(values)]
        return hash_strings(flat, self.lay.rows[f], self.device)

    def __call__(self, features, want_logx: bool = False, out: Optional[torch.Tensor] = None,
                 background: int = 0):
        """``background`` > 0: at most that many CTAs (ctr_criteo_rows_bg), for a copy stream beside a
        running step."""
        cont, cat = self.pack(features)
        B = (cont if cont is not None else cat).shape[0]
        rows = out if out is not None else \
            torch.empty((B, self.lay.F), dtype=torch.int32, device=self.device)
        if background > 0 and not want_logx and self.n_bnd <= 512:
            _call("ctr_criteo_rows_bg", _p(cont), len(self.cont_keys), _p(cat), len(self.cat_keys),
                  _p(self.fields_dev), _p(self.bnd_dev), self.n_bnd, B, self.lay.F, _p(rows),
                  _p(self.status), background, _stream())
            return rows
        logx = torch.empty((B, len(self.cont_keys)), dtype=torch.float32, device=self.device) \
            if want_logx else None
        _call("ctr_criteo_rows", _p(cont), len(self.cont_keys), _p(cat), len(self.cat_keys),
              _p(self.fields_dev), _p(self.bnd_dev), B, self.lay.F, _p(rows), _p(logx),
              _p(self.status), _stream())
        return (rows, logx) if want_logx else rows


def _is_int_array(v) -> bool:
    import numpy as np
    return isinstance(v, np.ndarray) and v.dtype.kind in "iu"


def hash_int64(ids: torch.Tensor, n_buckets: int, device) -> torch.Tensor:
    """Fingerprint64(as_string(id)) mod n_buckets on the device (ctr_hash_int64)."""
    ids = ids.to(device, torch.int64).contiguous()
    out = torch.empty_like(ids)
    _call("ctr_hash_int64", _p(ids), ids.numel(), int(n_buckets), _p(out), _stream())
    return out


def _flatten(values):
    for v in values:
        if isinstance(v, (bytes, bytearray)):
            yield v
        else:
            yield from _flatten(v)


def hash_strings(strings: Sequence[bytes], n_buckets: int, device) -> torch.Tensor:
    """FarmHash Fingerprint64(s) mod n_buckets on the device (ctr_hash_strings);
    TF StringToHashBucketFast as used by fm/fm.py:89."""
    n = len(strings)
    offs = [0]
    for s in strings:
        offs.append(offs[-1] + len(s))
    blob = torch.frombuffer(bytearray(b"".join(strings) + b"\0" * 16), dtype=torch.uint8).to(device)
    offsets = torch.tensor(offs, dtype=torch.int32, device=device)
    nb = torch.tensor([n_buckets], dtype=torch.int32, device=device)
    ro = torch.zeros(1, dtype=torch.int32, device=device)
    out = torch.empty(n, dtype=torch.int32, device=device)
    _call("ctr_hash_strings", _p(blob), _p(offsets), n, None, _p(nb), _p(ro), _p(out), _stream())
    return out.long()


# ---------------------------------------------------------------- Adam (TF rule)
class TFAdamState:
    """Step counter + lr_t of tf.train.AdamOptimizer (fm/fm.py:162):
    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), eps outside the sqrt.  The schedule lives on the
    device (``state`` = {t completed, lr_t of the step in progress, lr, counter}; see
    include/ctr_b200.h "optimiser") so a captured CUDA graph of the train step stays correct on
    replay.  It is advanced by the step's last optimiser launch (``DenseParams.adam_step``); a
    caller whose step has no dense parameters calls ``advance()`` itself."""

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, device=None):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.state = None
        if device is not None:
            self.state = torch.zeros(8, dtype=torch.float32, device=device)
            self.reset()

    def reset(self):
        """Back to t = 0 (tests; re-initialises the device schedule)."""
        self.t = 0
        if self.state is not None:
            self.state.zero_()
            self.state.view(torch.int32)[0] = -1       # t is a uint32 bit pattern: -1 + 1 = step 0
            self.advance()

    def advance(self):
        if self.state is not None:
            _call("ctr_adam_tick", _p(self.state), self.lr, self.beta1, self.beta2, _stream())

    def next_lr_t(self) -> float:
        """Host-side lr_t of the step that starts now (the device copy needs no launch: it was
        prepared when the previous step's optimiser finished)."""
        self.t += 1
        return self.lr * math.sqrt(1 - self.beta2 ** self.t) / (1 - self.beta1 ** self.t)

    @property
    def state_ptr(self):
        return None if self.state is None else self.state.data_ptr()


class DenseParams:
    """All dense (non-table) parameters of a model in ONE flat fp32 buffer, their
    gradients in another, so the optimiser is a single ctr_adam_dense launch.
    ``self[name]`` is a view with ``requires_grad`` whose ``.grad`` is a view of
    the flat gradient buffer."""

    def __init__(self, shapes: Dict[str, tuple], device, frozen: Sequence[str] = ()):
        self.names = list(shapes)
        self.shapes = dict(shapes)
        sizes = [int(torch.Size(s).numel()) for s in shapes.values()]
        # 4-float aligned slots so every view is 16-byte aligned
        self.slots, off = {}, 0
        for n, sz in zip(self.names, sizes):
            self.slots[n] = (off, sz)
            off += (sz + 3) // 4 * 4
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.frozen = set(frozen)
        # parameters whose 3xTF32 lo half was written by the latest optimiser launch (adam_step(lo=))
        # and is still valid; ``load`` - the other way parameters change - clears it
        self.lo_fresh = set()
        self.lo_targets: Dict[str, torch.Tensor] = {}
        self.views: Dict[str, torch.Tensor] = {}
        for n in self.names:
            o, sz = self.slots[n]
            v = self.flat[o:o + sz].view(self.shapes[n])
            if n not in self.frozen:
                v.requires_grad_(True)
                v.grad = self.grad[o:o + sz].view(self.shapes[n])
            self.views[n] = v

    def __getitem__(self, n):
        return self.views[n]

    def __contains__(self, n):
        return n in self.views

    def load(self, state: Dict[str, torch.Tensor]):
        self.lo_fresh.clear()
        with torch.no_grad():
            for n in self.names:
                if n in state:
                    self.views[n].copy_(state[n].to(self.flat.device, torch.float32)
                                        .reshape(self.shapes[n]))
        # registered lo halves are rebuilt right away: a step graph captured earlier contains no
        # split launch of its own
        for n, dst in self.lo_targets.items():
            if self.flat.is_cuda:
                _call("ctr_split_lo", _p(self.views[n]), _p(dst), dst.numel(), _stream())
                self.lo_fresh.add(n)

    def grads(self) -> Dict[str, torch.Tensor]:
        return {n: self.views[n].grad for n in self.names if n not in self.frozen}

    def zero_grad(self):
        self.grad.zero_()

    def adam_step(self, lr_t, st: TFAdamState, parties: int = 1, lo=None):
        """One launch over the flat buffer; frozen slots have zero gradient and
        zero moments, so the rule leaves them untouched.  ``parties``: optimiser kernels that end
        the step together (the last one to finish advances the device schedule).  ``lo`` =
        (name, dst): also write the 3xTF32 lo half of the updated parameter ``name`` to ``dst``."""
        lo_dst, lo_beg, lo_n = None, 0, 0
        self.lo_fresh.clear()
        if lo is not None:
            name, lo_dst = lo
            lo_beg, lo_n = self.slots[name]
            self.lo_fresh.add(name)
        _call("ctr_adam_dense_ex", _p(self.flat), _p(self.m), _p(self.v), _p(self.grad), self.numel,
              lr_t, st.beta1, st.beta2, st.eps, 1, st.state_ptr, parties, _p(lo_dst), lo_beg, lo_n,
              _stream())   # ends the step


# ------------------------------------------------------- fused multi-field lookup
class FieldEmbedding:
    """The concatenated embedding table [R, D] (+ first-order weights w1 [R]) of a
    list of embedding columns, with its gradient accumulators and Adam moments.

    Gradients are not autograd tensors: the backward kernel scatter-adds into the
    persistent ``dtable`` / ``dw1`` accumulators (dense, kept all-zero between steps
    by the optimiser kernels), which is what the reference's IndexedSlices
    gradient + AdamOptimizer amount to (fm/fm.py:162-163).
    """

    def __init__(self, lay: fc.Layout, device, with_w1: bool = True, w1_fields: int = 0,
                 adam_mode: str = "lazy", seed: int = 0, record: Optional[bool] = None):
        self.lay, self.device = lay, device
        self.D, self.F, self.R = lay.dimension, lay.F, lay.total_rows
        if self.D not in (8, 16, 32):
            raise ValueError("embedding_size must be 8, 16 or 32 (got %d)" % self.D)
        if self.F > 64:
            raise ValueError("at most 64 fields")
        D, R = self.D, self.R
        # Row-record layout (include/ctr_b200.h, "Row strides"): one [R, 4D+8] array holding
        # theta | m | v | g | theta1 m1 v1 g1 | claim per row, so that a lookup's first-order weight
        # and everything the lazy optimiser touches for a row share one DRAM page.  ``table``,
        # ``dtable``, ``w1`` ... are strided views of it.  The dense (exact_tf) optimiser streams
        # whole arrays and keeps the planar layout.
        if record is None:
            record = adam_mode == "lazy" and os.environ.get("CTR_ROW_RECORDS", "1") != "0"
        self.record = bool(record)
        g = torch.Generator(device=device).manual_seed(seed)
        if self.record:
            S = 4 * D + 8
            self.rec = torch.zeros(R, S, dtype=torch.float32, device=device)
            self.table, self._m = self.rec[:, 0:D], self.rec[:, D:2 * D]
            self._v, self.dtable = self.rec[:, 2 * D:3 * D], self.rec[:, 3 * D:4 * D]
            self.w1, self._m1 = self.rec[:, 4 * D], self.rec[:, 4 * D + 1]
            self._v1, self.dw1 = self.rec[:, 4 * D + 2], self.rec[:, 4 * D + 3]
            self._claim = self.rec[:, 4 * D + 4].view(torch.int32)
            self.ld = self.ld1 = self.ldc = S
        else:
            self.rec = None
            self.table = torch.empty(R, D, dtype=torch.float32, device=device)
            self.dtable = torch.zeros_like(self.table)
            self.w1 = torch.empty(R, dtype=torch.float32, device=device) if with_w1 else None
            self.dw1 = torch.zeros_like(self.w1) if with_w1 else None
            self._m = self._v = self._m1 = self._v1 = self._claim = None
            self.ld, self.ld1, self.ldc = D, 1, 1
        # embedding_column default initialiser: truncated normal, stddev 1/sqrt(D) [TF-sem];
        # drawn on the device (the R-full table is 2 GB), in row chunks so that the strided view
        # of a record array needs no table-sized temporary
        std = D ** -0.5
        step = 1 << 22
        for r0 in range(0, R, step):
            blk = torch.empty(min(step, R - r0), D, dtype=torch.float32, device=device)
            torch.nn.init.trunc_normal_(blk, std=std, a=-2 * std, b=2 * std, generator=g)
            self.table[r0:r0 + blk.shape[0]].copy_(blk)
        self.with_w1 = with_w1
        self.w1_fields = w1_fields
        if with_w1:
            lim = math.sqrt(6.0 / (R + 1))             # dense(onehot,1) glorot-uniform kernel
            self.w1.copy_((torch.rand(R, generator=g, device=device) * 2 - 1) * lim)
        elif self.record:
            self.w1 = self.dw1 = self._m1 = self._v1 = None
        self.adam_mode = adam_mode
        self._tag = 0
        self._offsets_host = (C.c_int64 * (self.F + 1))(*lay.offsets)
        self._anchor = torch.zeros((), device=device, requires_grad=True)
        # fused scatter + row optimiser (ctr_embed_bwd_adam): armed by ``arm_fused`` right before
        # the backward of a train step, consumed by the backward, acknowledged by ``adam_step``
        self.can_fuse = self.record and adam_mode == "lazy" and \
            os.environ.get("CTR_FUSED_ROW_ADAM", "1") != "0"
        self._fused = None
        self._fused_done = False
        self._side = None

    # -- state ------------------------------------------------------------------
    def load(self, table=None, w1=None):
        with torch.no_grad():
            if table is not None:
                self.table.copy_(table.to(self.device, torch.float32))
            if w1 is not None and self.with_w1:
                self.w1.copy_(w1.to(self.device, torch.float32).reshape(-1))

    def _ensure_adam(self):
        if self._m is None:
            self._m = torch.zeros_like(self.table)
            self._v = torch.zeros_like(self.table)
            self._claim = torch.zeros(self.R, dtype=torch.int32, device=self.device)
            if self.with_w1:
                self._m1 = torch.zeros_like(self.w1)
                self._v1 = torch.zeros_like(self.w1)

    # -- forward / backward -----------------------------------------------------
    def lookup(self, rows: torch.Tensor, want_fm: bool = True, want_y1: bool = True,
               cross_w: Optional[torch.Tensor] = None, cross_b: Optional[torch.Tensor] = None,
               want_lo: bool = False):
        """-> (E [B,F*D], y1 [B] (pre-bias), y2 [B], xl [B,F*D]) with autograd.  ``want_lo``: also
        write the lo half of E's 3xTF32 split (``self.last_E_lo``) for the tower's first GEMM."""
        require_cuda(rows, "rows")
        self._want_lo = bool(want_lo)
        return _EmbedFn.apply(self._anchor, self, rows, want_fm, want_y1 and self.with_w1,
                              cross_w, cross_b)

    def lookup_features(self, idp: "IdPipeline", features, want_logx: bool = False,
                        zero_buf: Optional[torch.Tensor] = None, tower0=None, **kw):
        """``lookup(idp(features))`` in ONE launch (ctr_embed_fwd_raw): the id pipeline runs as the
        first stage of the lookup kernel, which also clears ``zero_buf`` (the tower's per-step
        accumulators) on the side.  -> (rows, logx or None, lookup outputs...)."""
        if idp.n_bnd > 512 or os.environ.get("CTR_FUSED_IDS", "1") == "0":
            if zero_buf is not None:
                zero_buf.zero_()
            r = idp(features, want_logx=want_logx)
            rows, logx = r if want_logx else (r, None)
            return (rows, logx) + tuple(self.lookup(rows, **kw))
        cont, cat = idp.pack(features)
        B = (cont if cont is not None else cat).shape[0]
        pre = getattr(features, "rows", None)
        if pre is not None and not want_logx:
            # ids computed ahead of the step: the lookup takes them as they are (the plain kernel
            # stages them by TMA; the fused lookup + first layer kernel reads them directly)
            if tower0 is None:
                if zero_buf is not None:
                    zero_buf.zero_()
                return (pre, None) + tuple(self.lookup(pre, **kw))
            rows = pre
        else:
            pre = None
            rows = torch.empty((B, self.F), dtype=torch.int32, device=self.device)
        logx = torch.empty((B, len(idp.cont_keys)), dtype=torch.float32, device=self.device) \
            if want_logx else None
        self._raw = (idp, cont, cat, logx, zero_buf)
        self._rows_in = pre
        # tower0 = FusedTower: its first layer is computed by the lookup kernel itself
        # (ctr_embed_tower_fwd); the result is left in ``tower0.l0`` for ``tower_head``
        self._tower0 = tower0 if (tower0 is not None and not want_logx) else None
        try:
            outs = self.lookup(rows, **kw)
        finally:
            self._raw = None
            self._tower0 = None
        return (rows, logx) + tuple(outs)

    def zero_grad(self):
        self.dtable.zero_()
        if self.with_w1:
            self.dw1.zero_()

    def arm_fused(self, rows: torch.Tensor, lr_t: float, st: TFAdamState) -> bool:
        """Train step, right before the backward: count the batch's lookups per row
        (ctr_count_rows, on a side stream beside the tower's backward GEMMs) so that the
        backward can run scatter-add and Adam in one pass (ctr_embed_bwd_adam)."""
        if not self.can_fuse:
            return False
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        rows = rows.contiguous()
        with torch.cuda.stream(self._side):
            _call("ctr_count_rows", _p(rows), rows.numel(), self.D, _p(self.rec), self.ld,
                  self._side.cuda_stream)
            done = torch.cuda.Event()
            done.record(self._side)
        rows.record_stream(self._side)
        self._fused = (lr_t, st, done)
        return True

    def adam_step(self, rows: torch.Tensor, lr_t: float, st: TFAdamState, parties: int = 0):
        """Consume (and re-zero) the accumulated gradients.  ``lazy``: one update per
        touched row (LazyAdam); ``exact_tf``: every row, as TF's sparse apply does.  ``parties`` > 0:
        this launch is one of that many optimiser kernels ending the step together (lazy only)."""
        if self._fused_done:          # the backward already applied the update (ctr_embed_bwd_adam)
            self._fused_done = False
            return
        if self._fused is not None:   # armed, but no gradient reached the table: undo the counts
            raise RuntimeError("fused row optimiser was armed but the backward never reached the "
                               "embedding table")
        self._ensure_adam()
        if self.adam_mode == "exact_tf":
            _call("ctr_adam_dense", _p(self.table), _p(self._m), _p(self._v), _p(self.dtable),
                  self.table.numel(), lr_t, st.beta1, st.beta2, st.eps, 1, st.state_ptr, 0, _stream())
            if self.with_w1:
                _call("ctr_adam_dense", _p(self.w1), _p(self._m1), _p(self._v1), _p(self.dw1),
                      self.w1.numel(), lr_t, st.beta1, st.beta2, st.eps, 1, st.state_ptr, 0,
                      _stream())
            return
        self._tag += 1
        n = rows.numel()
        w = self.with_w1
        if rows.dim() == 2 and rows.shape[1] == self.F and rows.is_contiguous() and \
                os.environ.get("CTR_ADAM_ROWS_BF", "0") != "0":
            # the batch's [B, F] id matrix: the one-wave kernel (field-major warps, in-warp dedup)
            _call("ctr_adam_rows_bf", _p(rows), rows.shape[0], self.F, self.D, _p(self.table),
                  _p(self._m), _p(self._v), _p(self.dtable), _p(self.w1) if w else None,
                  _p(self._m1) if w else None, _p(self._v1) if w else None,
                  _p(self.dw1) if w else None, _p(self._claim), self._tag, lr_t, st.beta1, st.beta2,
                  st.eps, st.state_ptr, self.ld, self.ld1, self.ldc, parties, _stream())
            return
        _call("ctr_adam_rows_ex", _p(rows), n, self.D, _p(self.table), _p(self._m), _p(self._v),
              _p(self.dtable), _p(self.w1) if w else None, _p(self._m1) if w else None,
              _p(self._v1) if w else None, _p(self.dw1) if w else None, _p(self._claim), self._tag,
              lr_t, st.beta1, st.beta2, st.eps, st.state_ptr, self.ld, self.ld1, self.ldc, parties,
              _stream())


class _EmbedFn(torch.autograd.Function):
    """ctr_embed_fwd / ctr_embed_bwd (+ ctr_dcn_cross_bwd when the cross stack is fused)."""

    @staticmethod
    def forward(ctx, anchor, emb: FieldEmbedding, rows, want_fm, want_y1, cross_w, cross_b):
        ctx.set_materialize_grads(False)
        B, F, D = rows.shape[0], emb.F, emb.D
        dev = rows.device
        rows = rows.contiguous()
        ctx.tw0 = None
        E = torch.empty((B, F * D), dtype=torch.float32, device=dev)
        S = torch.empty((B, D), dtype=torch.float32, device=dev) if want_fm else None
        y2 = torch.empty((B,), dtype=torch.float32, device=dev) if want_fm else None
        y1 = torch.empty((B,), dtype=torch.float32, device=dev) if want_y1 else None
        cross = cross_w is not None
        xl = torch.empty((B, F * D), dtype=torch.float32, device=dev) if cross else None
        L = cross_w.shape[0] if cross else 0
        E_lo = torch.empty_like(E) if getattr(emb, "_want_lo", False) else None
        emb.last_E_lo = E_lo
        raw = getattr(emb, "_raw", None)
        tw0 = getattr(emb, "_tower0", None) if raw is not None else None
        if tw0 is not None and not cross and E_lo is not None:
            # ids + gather + FM terms + first tower layer in one launch
            idp, cont, cat, logx, zbuf = raw
            N = tw0.sizes[1]
            act0 = torch.empty((B, N), dtype=torch.float32, device=dev)
            nparts = (B + 127) // 128
            parts = torch.empty((nparts, 2, N), dtype=torch.float32, device=dev)
            tw0.wait_w0_lo()
            _call("ctr_embed_tower_fwd", _p(emb.table), _p(emb.w1), _p(cont), len(idp.cont_keys),
                  _p(cat), len(idp.cat_keys), _p(idp.fields_dev), _p(idp.bnd_dev), idp.n_bnd,
                  _p(getattr(emb, "_rows_in", None)), _p(rows),
                  _p(idp.status), B, F, D, emb.w1_fields, _p(E), _p(E_lo), _p(S), _p(y1), _p(y2),
                  emb.ld, emb.ld1, _p(tw0.P("0.w")), _p(tw0.w0_lo), _p(tw0.P("0.b")), N, _p(act0),
                  _p(parts), _p(zbuf), zbuf.numel() if zbuf is not None else 0, _stream())
            tw0.l0 = (act0, parts)
            ctx.tw0 = tw0
        elif raw is not None:       # id pipeline fused in front: fills ``rows`` (and logx) as well
            idp, cont, cat, logx, zbuf = raw
            _call("ctr_embed_fwd_raw", _p(emb.table), _p(emb.w1), _p(cont), len(idp.cont_keys),
                  _p(cat), len(idp.cat_keys), _p(idp.fields_dev), _p(idp.bnd_dev), idp.n_bnd,
                  _p(rows), _p(logx), _p(idp.status), B, F, D, emb.w1_fields, _p(E), _p(S), _p(y1),
                  _p(y2), _p(cross_w), _p(cross_b), L, _p(xl), _p(E_lo), emb.ld, emb.ld1, _p(zbuf),
                  zbuf.numel() if zbuf is not None else 0, _stream())
        else:
            _call("ctr_embed_fwd", _p(emb.table), _p(emb.w1), _p(rows), B, F, D, emb.w1_fields, _p(E),
                  _p(S), _p(y1), _p(y2), _p(cross_w), _p(cross_b), L, _p(xl), _p(E_lo), emb.ld,
                  emb.ld1, _stream())
        ctx.emb, ctx.rows, ctx.E, ctx.S = emb, rows, E, S
        ctx.cross_w, ctx.cross_b = cross_w, cross_b
        ctx.flags = (want_fm, want_y1, cross)
        # placeholders for the outputs not asked for: uninitialised scalars (no fill launch)
        outs = [E, y1 if want_y1 else E.new_empty(()), y2 if want_fm else E.new_empty(()),
                xl if cross else E.new_empty(())]
        non_diff = [o for o, f in zip(outs[1:], (want_y1, want_fm, cross)) if not f]
        if non_diff:
            ctx.mark_non_differentiable(*non_diff)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dE, dy1, dy2, dxl):
        emb = ctx.emb
        want_fm, want_y1, cross = ctx.flags
        rows, E, S = ctx.rows, ctx.E, ctx.S
        B, F, D = rows.shape[0], emb.F, emb.D
        dcw = dcb = None
        if cross and dxl is not None:
            L, W = ctx.cross_w.shape
            dx0 = torch.empty_like(E)
            dcw = torch.zeros_like(ctx.cross_w)
            dcb = torch.zeros_like(ctx.cross_b)
            _call("ctr_dcn_cross_bwd", _p(E), _p(ctx.cross_w), _p(ctx.cross_b), L, B, W,
                  _p(dxl.contiguous()), _p(dx0), _p(dcw), _p(dcb), _stream())
            dE = dx0 if dE is None else dE + dx0
        dy2 = dy2.contiguous() if (want_fm and dy2 is not None) else None
        dy1 = dy1.contiguous() if (want_y1 and dy1 is not None) else None
        tw0 = ctx.tw0
        bwd0 = getattr(tw0, "bwd0", None) if tw0 is not None else None
        if bwd0 is not None:
            # the tower left dpre0 instead of dE: first-layer data gradient + scatter-add in one
            # launch (ctr_tower_embed_bwd); ``dE`` is a placeholder
            tw0.bwd0 = None
            dpre0, dpre0_lo, after_scatter = bwd0
            if emb._fused is not None:      # one-pass scatter + optimiser armed: it wants dE itself
                dE = torch.empty((B, F * D), dtype=torch.float32, device=rows.device)
                _call("ctr_tower_gemm_presplit", 1, _p(dpre0), _p(dpre0_lo), _p(tw0.P("0.w")),
                      _p(tw0.w0_lo), B, F * D, tw0.sizes[1], _p(dE), None, None, 0, _stream())
                after_scatter()
                bwd0 = None
        if bwd0 is not None:
            _call("ctr_tower_embed_bwd", _p(dpre0), _p(dpre0_lo), _p(tw0.P("0.w")), _p(tw0.w0_lo),
                  tw0.sizes[1], _p(rows), _p(E), _p(S), _p(dy2), _p(dy1), emb.w1_fields,
                  emb._offsets_host, B, F, D, _p(emb.dtable), _p(emb.dw1), emb.ld, emb.ld1, _stream())
            after_scatter()
            return None, None, None, None, None, dcw, dcb
        dE = None if dE is None else dE.contiguous()
        fused, emb._fused = emb._fused, None
        if fused is not None and (dE is not None or dy2 is not None):
            lr_t, st, counted = fused
            torch.cuda.current_stream().wait_event(counted)
            _call("ctr_embed_bwd_adam", _p(rows), _p(dE), _p(S), _p(dy2), _p(dy1), emb.w1_fields,
                  emb._offsets_host, B, F, D, _p(emb.rec), emb.ld, lr_t, st.beta1, st.beta2, st.eps,
                  st.state_ptr, _stream())
            emb._fused_done = True
        elif dE is not None or dy2 is not None:
            _call("ctr_embed_bwd", _p(rows), _p(dE), _p(E), _p(emb.table), _p(S), _p(dy2), _p(dy1),
                  emb.w1_fields, emb._offsets_host, B, F, D, _p(emb.dtable), _p(emb.dw1), emb.ld,
                  emb.ld1, _stream())
        return None, None, None, None, None, dcw, dcb


# ------------------------------------------------------------------- DCN cross
class _CrossFn(torch.autograd.Function):
    """Stand-alone cross stack on a dense x0 (ctr_dcn_cross_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x0, w, b):
        x0, w, b = x0.contiguous(), w.contiguous(), b.contiguous()
        L, W = w.shape
        xl = torch.empty_like(x0)
        _call("ctr_dcn_cross_fwd", _p(x0), _p(w), _p(b), L, x0.shape[0], W, _p(xl), _stream())
        ctx.save_for_backward(x0, w, b)
        return xl

    @staticmethod
    def backward(ctx, dxl):
        x0, w, b = ctx.saved_tensors
        L, W = w.shape
        dx0 = torch.empty_like(x0)
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        _call("ctr_dcn_cross_bwd", _p(x0), _p(w), _p(b), L, x0.shape[0], W, _p(dxl.contiguous()),
              _p(dx0), _p(dw), _p(db), _stream())
        return dx0, dw, db


def dcn_cross(x0, w, b):
    require_cuda(x0, "x0")
    return _CrossFn.apply(x0, w, b)


# ------------------------------------------------------------------ xDeepFM CIN
CIN_PREC = {"fp32": 0, "tf32": 1, "tf32x3": 2}
_WS_CACHE: Dict[tuple, torch.Tensor] = {}


def _workspace(nbytes: int, device, tag: str) -> torch.Tensor:
    key = (tag, str(device))
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


class _CINFn(torch.autograd.Function):
    """The whole CIN stack (xdeepfm/xdeepfm.py:135-181): E [B, m*D] -> pooled [B, sum H].
    Layers run through ctr_cin_layer_fwd/_bwd on d-major rows; the [B,D,m*Hp] outer
    product the reference materialises never exists."""

    @staticmethod
    def forward(ctx, E, m, D, prec, *wb):
        n = len(wb) // 2
        Ws, bs = wb[:n], wb[n:]
        B = E.shape[0]
        dev = E.device
        ld0 = (m + 3) // 4 * 4
        X0t = torch.empty((B * D, ld0), dtype=torch.float32, device=dev)
        _call("ctr_transpose_fd", _p(E.contiguous()), B, m, D, _p(X0t), ld0, _stream())
        lib = _lib.load()
        hp_list = [m] + [int(W.shape[1]) for W in Ws[:-1]]
        nbytes = max(int(lib.ctr_cin_workspace_bytes(B, D, m, hp, int(W.shape[1]), prec))
                     for hp, W in zip(hp_list, Ws))
        ws = _workspace(nbytes, dev, "cin")
        outs = []
        Xp, ldp, Hp = X0t, ld0, m
        for W, b in zip(Ws, bs):
            H = int(W.shape[1])
            out = torch.empty((B * D, H), dtype=torch.float32, device=dev)
            _call("ctr_cin_layer_fwd", _p(X0t), ld0, _p(Xp), ldp, _p(W.contiguous()), _p(b), B, D,
                  m, Hp, H, _p(out), prec, _p(ws), ws.numel(), _stream())
            outs.append(out)
            Xp, ldp, Hp = out, H, H
        Hs = [int(W.shape[1]) for W in Ws]
        if all(h % 4 == 0 for h in Hs):       # sum over d straight into the concatenated tensor
            pooled = torch.empty((B, sum(Hs)), dtype=torch.float32, device=dev)
            off = 0
            for o, h in zip(outs, Hs):
                _call("ctr_cin_pool", _p(o), B, D, h, _p(pooled) + 4 * off, pooled.shape[1], _stream())
                off += h
        else:
            pooled = torch.cat([o.view(B, D, -1).sum(1) for o in outs], 1)      # :180-181
        ctx.meta = (m, D, prec, ld0, B)
        ctx.saved = (X0t, outs, Ws, ws)
        return pooled

    @staticmethod
    def backward(ctx, dp):
        m, D, prec, ld0, B = ctx.meta
        X0t, outs, Ws, ws = ctx.saved
        dev = dp.device
        n = len(Ws)
        Hs = [int(W.shape[1]) for W in Ws]
        offs = [0]
        for h in Hs:
            offs.append(offs[-1] + h)
        dX0t = torch.zeros_like(X0t)
        dWs = [torch.zeros_like(W) for W in Ws]
        dbs = [torch.zeros(h, dtype=torch.float32, device=dev) for h in Hs]
        fused = all(h % 4 == 0 for h in Hs) and dp.shape[1] % 4 == 0
        dp = dp.contiguous()
        if fused:
            # dacc[k]: what layer k+1's backward accumulates for layer k's output (zero for the last)
            dacc = [torch.zeros((B * D, Hs[k]), dtype=torch.float32, device=dev) if k < n - 1 else None
                    for k in range(n)]
        else:
            # gradient arriving at each layer output from the sum-pool (broadcast over d)
            douts = [dp[:, offs[k]:offs[k + 1]].unsqueeze(1).expand(B, D, Hs[k]).reshape(B * D, Hs[k])
                     .contiguous() for k in range(n)]
        for k in range(n - 1, -1, -1):
            if fused:   # (pool gradient broadcast over d + dacc) through the ReLU, one kernel
                dpre = torch.empty((B * D, Hs[k]), dtype=torch.float32, device=dev)
                _call("ctr_cin_dpre", _p(dp) + 4 * offs[k], dp.shape[1], _p(dacc[k]), _p(outs[k]), B, D,
                      Hs[k], _p(dpre), _stream())
            else:
                dpre = (douts[k] * (outs[k] > 0)).contiguous()
            if k > 0:
                Xp, ldp, Hp, dXp = outs[k - 1], Hs[k - 1], Hs[k - 1], (dacc if fused else douts)[k - 1]
            else:
                Xp, ldp, Hp, dXp = X0t, ld0, m, dX0t
            _call("ctr_cin_layer_bwd", _p(X0t), ld0, _p(Xp), ldp, _p(Ws[k].contiguous()), _p(dpre),
                  B, D, m, Hp, Hs[k], _p(dX0t), _p(dXp), _p(dWs[k]), _p(dbs[k]), prec, _p(ws),
                  ws.numel(), _stream())
        dE = torch.zeros((B, m * D), dtype=torch.float32, device=dev)
        _call("ctr_transpose_df_add", _p(dX0t), ld0, B, m, D, _p(dE), _stream())
        return (dE, None, None, None) + tuple(dWs) + tuple(dbs)


def cin(E, m, D, Ws, bs, precision="tf32x3"):
    require_cuda(E, "E")
    return _CINFn.apply(E, m, D, CIN_PREC[precision], *Ws, *bs)


# ------------------------------------------------------------ DIN activation unit
class _DinAttFn(torch.autograd.Function):
    """din/din.py:103-125 through ctr_din_att_fwd / ctr_din_att_bwd.  ``table`` /
    ``dtable`` are views of a FieldEmbedding's buffers starting at the sub-table's
    first row; the table gradient is RED-accumulated there, not returned."""

    @staticmethod
    def forward(ctx, anchor, table, dtable, hist, query, W1, b1, W2, b2, W3, b3, opts):
        B, P = hist.shape
        E = query.shape[1]
        hist = hist.contiguous()
        query = query.contiguous()
        out = torch.empty((B, E), dtype=torch.float32, device=query.device)
        _call("ctr_din_att_fwd", _p(table), _p(hist), _p(query), B, P, E, _p(W1), _p(b1),
              W1.shape[1], _p(W2), _p(b2), W2.shape[1], _p(W3), _p(b3), _p(out), None,
              C.byref(opts) if opts is not None else None, _stream())
        ctx.save_for_backward(hist, query, W1, b1, W2, b2, W3, b3)
        ctx.tabs = (table, dtable, opts)
        return out

    @staticmethod
    def backward(ctx, dout):
        hist, query, W1, b1, W2, b2, W3, b3 = ctx.saved_tensors
        table, dtable, opts = ctx.tabs
        B, P = hist.shape
        E = query.shape[1]
        lib = _lib.load()
        nbytes = int(lib.ctr_din_workspace_bytes(B, P, E))
        ws = _workspace(nbytes, query.device, "din")
        dq = torch.empty_like(query)
        g = [torch.zeros_like(t) for t in (W1, b1, W2, b2, W3, b3)]
        _call("ctr_din_att_bwd", _p(table), _p(hist), _p(query), B, P, E, _p(W1), _p(b1),
              W1.shape[1], _p(W2), _p(b2), W2.shape[1], _p(W3), _p(b3), _p(dout.contiguous()),
              _p(dtable), _p(dq), _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]), _p(g[5]),
              _p(ws), ws.numel(), C.byref(opts) if opts is not None else None, _stream())
        return (None, None, None, None, dq) + tuple(g) + (None,)


def din_opts(unit: int, n_rows: int, p_drop: float = 0.0, seed: int = 0, state_ptr=None,
             status: Optional[torch.Tensor] = None) -> "_lib.DinOpts":
    """ctr_din_opts: dropout stream + id range check of one attention unit."""
    o = _lib.DinOpts()
    o.state, o.p_drop, o.seed, o.unit = state_ptr, float(p_drop), seed & 0xFFFFFFFF, unit
    o.table_rows, o.status = int(n_rows), _p(status)
    return o


def din_attention(emb: "FieldEmbedding", field: int, hist, query, W1, b1, W2, b2, W3, b3,
                  opts=None):
    """Attention of ``query`` over the history ids ``hist`` (local ids of sub-table
    ``field`` of ``emb``; 0 = padding).  ``opts``: ``din_opts(...)`` (dropout, id range check)."""
    require_cuda(hist, "hist")
    lo, hi = emb.lay.offsets[field], emb.lay.offsets[field + 1]
    if opts is None:
        opts = din_opts(field, hi - lo)
    return _DinAttFn.apply(emb._anchor, emb.table[lo:hi], emb.dtable[lo:hi], hist, query,
                           W1.contiguous(), b1, W2.contiguous(), b2, W3.reshape(-1), b3, opts)


def din_dropout_masks(opts, n_rows: int, device):
    """Keep scales (0 or 1/(1-p)) the attention kernels use at the current step:
    ([n_rows, 80], [n_rows, 40]) - what the parity tests inject into the oracle."""
    outs = []
    for layer, H in ((0, 80), (1, 40)):
        o = torch.empty((n_rows, H), dtype=torch.float32, device=device)
        _call("ctr_din_dropout_mask", C.byref(opts), layer, n_rows, H, _p(o), _stream())
        outs.append(o)
    return outs


# ------------------------------------------------------- fused dense tower + loss head
BN_EPS = 1e-3


class FusedTower:
    """dense(relu) -> BN -> dropout stack (+ optional final dense(1, relu)) through the
    ctr_tower_* kernels (deepfm/deepfm.py:100-108).  Parameters live in ``dense``
    (DenseParams) under ``prefix``; weight / BN gradients are accumulated straight into the
    flat gradient buffer by the backward kernels (no per-parameter autograd adds)."""

    def __init__(self, dense: "DenseParams", prefix: str, sizes, out_layer: bool, dropout: float,
                 adam: "TFAdamState", seed: int = 0, bn: bool = True, out_relu: bool = True,
                 layer_base: int = 0):
        """``bn=False``: dense(relu) -> dropout without batch normalisation (DIN's MLP,
        din/din.py:133-137).  ``out_relu=False``: the final dense(1) has no activation
        (din/din.py:139).  ``layer_base``: offset of this tower's layers in the dropout
        counter space, so that two towers of one model draw independent masks."""
        self.dense, self.prefix, self.sizes = dense, prefix, list(sizes)
        self.out_layer, self.dropout, self.adam, self.seed = out_layer, float(dropout), adam, seed
        self.bn, self.out_relu, self.layer_base = bool(bn), bool(out_relu), int(layer_base)
        if not self.bn:     # identity "BN" constants for the prologue descriptors
            w = max(self.sizes)
            self._zeros = torch.zeros(w, dtype=torch.float32, device=dense.flat.device)
            self._ones = torch.ones(w, dtype=torch.float32, device=dense.flat.device)
        self._ones_col = {}
        self._anchor = torch.zeros((), device=dense.flat.device, requires_grad=True)
        # weight-gradient kernels run on a side stream, off the critical path
        # (dX chain -> embedding scatter -> row Adam); joined in ``join()`` before dense Adam
        self.side = torch.cuda.Stream(device=dense.flat.device)
        self._pending = None
        # write each hidden layer's dpre once (ctr_tower_dpre) instead of re-deriving it from
        # the BN bookkeeping inside both backward GEMMs; required for the tcgen05 path
        self.materialise_dpre = os.environ.get("CTR_TOWER_DPRE", "1") != "0"
        # ctr_tower_mid: the middle of the chain (hidden layers >= 1, dense(1), head, loss and
        # their backward) in one cooperative launch; needs every hidden width <= 128
        self.barrier = torch.zeros(320, dtype=torch.int32, device=dense.flat.device)   # CTR_TOWER_MID_BARRIER_WORDS
        self.timing = None      # set to an int64[8] device tensor to get the phase time stamps
        self.last_acts = None
        self.last_y = None
        self.mid_ok = (out_layer and self.bn and self.out_relu and 1 <= len(self.sizes) - 1 <= 4
                       and all(4 <= h <= 128 and h % 4 == 0 for h in self.sizes[1:]))

        # pre-split 3xTF32 operands for the first layer's tcgen05 GEMMs (ctr_tower_gemm_presplit)
        self.w0_lo = torch.zeros(self.sizes[0] * self.sizes[1], dtype=torch.float32,
                                 device=dense.flat.device)
        self._w0_ready = None
        self._w0_fresh = False
        dense.lo_targets[prefix + ".0.w"] = self.w0_lo
        self._ws = None
        self.l0 = None          # (act0, stats_part) left by the fused lookup + first layer kernel
        self.bwd0 = None        # (dpre0, dpre0_lo) left for the fused data-gradient + scatter kernel
        self._zero = torch.zeros((), dtype=torch.float32, device=dense.flat.device)
        # default off: measured slower than GEMM + scatter at batch 4096 (profiles/r02)
        self.fuse_bwd0 = os.environ.get("CTR_FUSED_BWD0", "0") != "0" and self.sizes[1] <= 128

    def can_fuse_l0(self, F: int, D: int) -> bool:
        """The first layer can be computed inside the lookup kernel (ctr_embed_tower_fwd)."""
        return (self.use_presplit and D == 16 and 8 < F <= 40 and self.sizes[0] == F * D
                and 16 <= self.sizes[1] <= 128 and os.environ.get("CTR_FUSED_L0", "1") != "0")

    def wait_w0_lo(self):
        """The current stream waits for the hi/lo split of the first layer's weights."""
        if self._w0_ready is None:
            self._begin_split()
        if self._w0_ready is not None:
            torch.cuda.current_stream().wait_event(self._w0_ready)
            self._w0_ready = None

    @property
    def use_mid(self):
        return self.mid_ok and os.environ.get("CTR_TOWER_MID", "1") != "0"

    @property
    def use_presplit(self):
        return (self.use_mid and os.environ.get("CTR_TOWER_PRESPLIT", "1") != "0"
                and self.sizes[0] % 4 == 0 and self.sizes[0] >= 32 and self.sizes[1] >= 16)

    def ws_numel(self, B, splitk):
        """Floats of the per-step workspace: BN column sums of every layer | loss (+3 pad) |
        (split-K first GEMM) its [B, H0] accumulation target."""
        return 2 * sum(self.sizes[1:]) + 4 + (B * self.sizes[1] if splitk else 0)

    def splitk_for(self, B, have_lo):
        return bool(have_lo) and self.use_presplit and B >= 256 and \
            os.environ.get("CTR_TOWER_SPLITK", "1") != "0"

    def begin_step(self, B=None, expect_lo=False, fused_l0=False):
        """Start-of-step hook.  (1) Split the first layer's weights into hi/lo on the side stream
        (they only change in the optimiser), off the critical path of the id/lookup kernels.
        (2) With ``B``: allocate the step's workspace and return it, so that the caller can have
        the lookup kernel clear it (``FieldEmbedding.lookup_features(zero_buf=...)``); the tower
        then takes it as already zeroed."""
        self._ws = None
        self.l0 = None
        ws = None
        if B is not None and self.use_mid:
            splitk = self.splitk_for(B, expect_lo) and not fused_l0
            ws = torch.empty(self.ws_numel(B, splitk), dtype=torch.float32,
                             device=self.dense.flat.device)
            self._ws = (ws, B, splitk)
        self._begin_split()
        return ws

    def _begin_split(self):
        if not self.use_presplit:
            return
        if (self.prefix + ".0.w") in self.dense.lo_fresh:
            # the optimiser launch that updated W0 wrote its lo half as well (ordered before
            # everything of this step by the end-of-step join): nothing to launch, nothing to wait for
            self._w0_ready = None
            self._w0_fresh = True
            return
        self._w0_fresh = False
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            _call("ctr_split_lo", _p(self.P("0.w")), _p(self.w0_lo), self.w0_lo.numel(),
                  self.side.cuda_stream)
            done = torch.cuda.Event()
            done.record(self.side)
        self._w0_ready = done

    def join(self):
        """Make the current stream wait for the side-stream weight-gradient kernels."""
        if self._pending is not None:
            torch.cuda.current_stream().wait_event(self._pending)
            self._pending = None

    def P(self, name):
        return self.dense[self.prefix + "." + name]

    def G(self, name):
        key = self.prefix + "." + name
        return self.dense[key].grad if key in self.dense else None

    def ones_col(self, B):
        t = self._ones_col.get(B)
        if t is None:
            t = self._ones_col[B] = torch.ones(B, 1, dtype=torch.float32,
                                               device=self.dense.flat.device)
        return t

    def __call__(self, X, training: bool):
        require_cuda(X, "tower input")
        return _TowerFn.apply(X, self._anchor, self, training)

    def bn_drop(self, l, sums, training):
        d = _lib.BnDrop()
        d.state = self.adam.state_ptr
        d.p_drop = self.dropout if training else 0.0
        d.seed, d.layer = self.seed & 0xFFFFFFFF, self.layer_base + l
        if self.bn:
            d.sums = _p(sums) if training else None
            d.mean, d.var = _p(self.P("%d.bn.mean" % l)), _p(self.P("%d.bn.var" % l))
            d.gamma, d.beta = _p(self.P("%d.bn.gamma" % l)), _p(self.P("%d.bn.beta" % l))
            d.eps, d.enabled = BN_EPS, 1
        else:       # x' = x * keep: mean 0, var 1, eps 0, gamma 1, beta 0; nothing at all in eval
            d.sums = None
            d.mean, d.var = _p(self._zeros), _p(self._ones)
            d.gamma, d.beta = _p(self._ones), _p(self._zeros)
            d.eps, d.enabled = 0.0, 1 if d.p_drop > 0.0 else 0
        return d


class _TowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, anchor, tw: FusedTower, training):
        ctx.set_materialize_grads(False)
        X = X.contiguous()
        B, dev = X.shape[0], X.device
        L = len(tw.sizes) - 1
        acts, stats, pros = [], [], []
        x, K, pro = X, tw.sizes[0], None
        for l in range(L):
            H = tw.sizes[l + 1]
            a = torch.empty((B, H), dtype=torch.float32, device=dev)
            st = torch.zeros((2, H), dtype=torch.float32, device=dev) if (training and tw.bn) else None
            _call("ctr_tower_layer_fwd", _p(x), K, K, C.byref(pro) if pro is not None else None,
                  _p(tw.P("%d.w" % l)), _p(tw.P("%d.b" % l)), H, _p(a), H, _p(st), 1, B, _stream())
            pro = tw.bn_drop(l, st, training)
            acts.append(a)
            stats.append(st)
            pros.append(pro)
            x, K = a, H
        if tw.out_layer:
            y = torch.empty((B, 1), dtype=torch.float32, device=dev)
            _call("ctr_tower_layer_fwd", _p(x), K, K, C.byref(pro), _p(tw.P("out.w")),
                  _p(tw.P("out.b")), 1, _p(y), 1, None, 1 if tw.out_relu else 0, B, _stream())
            out = y.view(B)
        elif pro.enabled:
            y = None
            out = torch.empty((B, K), dtype=torch.float32, device=dev)
            _call("ctr_bn_drop_apply", _p(x), K, C.byref(pro), _p(out), B, _stream())
        else:
            y, out = None, x.clone()
        ctx.tw, ctx.training = tw, training
        ctx.saved = (X, acts, stats, pros, y)
        return out

    @staticmethod
    def backward(ctx, dout):
        tw = ctx.tw
        X, acts, stats, pros, y = ctx.saved
        if not ctx.training:
            raise RuntimeError("FusedTower backward is only defined in training mode")
        B, dev = X.shape[0], X.device
        L = len(tw.sizes) - 1
        dout = dout.contiguous().view(B, -1)

        def grad_src(G, ldg, a, lda, kind, l=None):
            if kind == 1 and not tw.bn:
                kind = 0          # no BN after the layer: the stored dn is the gradient itself
            g = _lib.GradSrc()
            g.G, g.ldg, g.a, g.lda, g.kind, g.train, g.eps = _p(G), ldg, _p(a), lda, kind, 1, BN_EPS
            if kind == 1:
                g.sums = _p(stats[l])
                g.gamma = _p(tw.P("%d.bn.gamma" % l))
                g.dbeta, g.dgamma = _p(tw.G("%d.bn.beta" % l)), _p(tw.G("%d.bn.gamma" % l))
            return g

        main = torch.cuda.current_stream()
        side = tw.side

        def fork():
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)

        def weights(xin, K, pro, gs_w, H, wname, bname):
            with torch.cuda.stream(side):
                _call("ctr_tower_layer_bwd_weights", _p(xin), K, K, pro, C.byref(gs_w), H,
                      _p(tw.G(wname)), _p(tw.G(bname)) if bname else None, B, side.cuda_stream)

        HL = tw.sizes[L]
        dn = torch.empty((B, HL), dtype=torch.float32, device=dev)
        if tw.out_layer:
            # final dense(1, relu): dpre = dout * 1[y > 0]; without the ReLU (DIN) dpre = dout, which
            # the same kernels produce when the "activation" they gate on is a column of ones
            gs = grad_src(dout, 1, y if tw.out_relu else tw.ones_col(B), 1, 0)
            fork()
            weights(acts[L - 1], HL, C.byref(pros[L - 1]), gs, 1, "out.w", "out.b")
            _call("ctr_tower_layer_bwd_data", C.byref(gs), 1, _p(tw.P("out.w")), HL,
                  C.byref(pros[L - 1]), _p(acts[L - 1]), _p(dn), HL,
                  _p(tw.G("%d.bn.beta" % (L - 1))), _p(tw.G("%d.bn.gamma" % (L - 1))), B, _stream())
        elif pros[L - 1].enabled:
            # the tower ends in BN + dropout (dcn/dcn.py:146-149): through the dropout, and the BN
            # column sums of the last layer
            _call("ctr_bn_drop_apply_bwd", _p(dout), HL, _p(acts[L - 1]), HL, C.byref(pros[L - 1]),
                  _p(dn), _p(tw.G("%d.bn.beta" % (L - 1))), _p(tw.G("%d.bn.gamma" % (L - 1))), B,
                  _stream())
        else:
            dn = dout
        keep = [dout, dn]
        for l in range(L - 1, -1, -1):
            H, K = tw.sizes[l + 1], tw.sizes[l]
            gs = grad_src(dn, H, acts[l], H, 1, l)
            xin = acts[l - 1] if l > 0 else X
            pro = C.byref(pros[l - 1]) if l > 0 else None
            bname = "%d.b" % l
            if tw.materialise_dpre:
                # dpre_l written once (+ db_l), then both GEMMs read it as a plain tensor
                dpre = torch.empty((B, H), dtype=torch.float32, device=dev)
                _call("ctr_tower_dpre", C.byref(gs), H, _p(dpre), H, _p(tw.G(bname)), B, _stream())
                gs = grad_src(dpre, H, None, 0, 2)
                keep.append(dpre)
                bname = None
            fork()        # dbeta_l / dgamma_l (written by the data kernel above) are ready
            weights(xin, K, pro, gs, H, "%d.w" % l, bname)
            dnext = torch.empty((B, K), dtype=torch.float32, device=dev)
            _call("ctr_tower_layer_bwd_data", C.byref(gs), H, _p(tw.P("%d.w" % l)), K, pro,
                  _p(xin) if l > 0 else None, _p(dnext), K,
                  _p(tw.G("%d.bn.beta" % (l - 1))) if l > 0 else None,
                  _p(tw.G("%d.bn.gamma" % (l - 1))) if l > 0 else None, B, _stream())
            dn = dnext
            keep.append(dn)
        for t_ in keep + acts + [X]:           # tensors read by the side stream
            t_.record_stream(side)
        done = torch.cuda.Event()
        done.record(side)
        tw._pending = done
        return dn, None, None, None


class _TowerHeadFn(torch.autograd.Function):
    """Tower + loss head with the middle of the chain in ONE cooperative launch (ctr_tower_mid):
    layer-0 GEMM -> [hidden layers, dense(1), head, loss, and in training the whole backward down
    to dpre_0] -> (backward) dX GEMM on the main stream, weight-gradient GEMMs on the side stream.
    All parameter gradients are accumulated into the flat gradient buffer by the kernels."""

    @staticmethod
    def forward(ctx, X, anchor, tw: "FusedTower", head, labels, training, X_lo, *zs):
        ctx.set_materialize_grads(False)
        dense = tw.dense
        hw, hb, b1, relu0, grad_scale = head
        X = X.contiguous()
        B, dev = X.shape[0], X.device
        L = len(tw.sizes) - 1
        Hs = tw.sizes[1:]
        zs = [z.contiguous().view(B) for z in zs]
        labels = labels.to(dev, torch.float32).contiguous().view(B)
        f32 = dict(dtype=torch.float32, device=dev)
        # one zero-filled workspace per step: the BN column sums of every layer + the loss scalar
        offs, n = [], 0
        for H in Hs:
            offs.append(n)
            n += 2 * H
        presplit = X_lo is not None and tw.use_presplit and B >= 256
        l0, tw.l0 = tw.l0, None
        if l0 is not None and (l0[0].shape[0] != B or not presplit):
            l0 = None
        splitk = tw.splitk_for(B, X_lo is not None) and l0 is None
        # (+ with the split-K first GEMM: its zero-initialised accumulation target).  Taken from
        # begin_step() when the lookup kernel has already cleared it, else one fill launch here.
        pre, tw._ws = tw._ws, None
        if pre is not None and pre[1] == B and pre[2] == splitk and pre[0].numel() == tw.ws_numel(B, splitk):
            ws = pre[0]
        else:
            ws = torch.zeros(tw.ws_numel(B, splitk), **f32)
        stats = [ws[o:o + 2 * H].view(2, H) for o, H in zip(offs, Hs)]
        loss = ws[n:n + 1].view(())
        pre0 = ws[n + 4:].view(B, Hs[0]) if splitk else None
        acts = [torch.empty((B, H), **f32) for H in Hs]
        logits, prob, y = (torch.empty(B, **f32) for _ in range(3))
        if l0 is not None:          # the lookup kernel has already produced act0 and its column sums
            acts[0] = l0[0]
        elif presplit:
            tw.wait_w0_lo()
            if splitk:      # partial sums only; bias / ReLU / column sums happen in ctr_tower_mid
                _call("ctr_tower_gemm_presplit", 3, _p(X), _p(X_lo), _p(tw.P("0.w")), _p(tw.w0_lo),
                      B, tw.sizes[0], Hs[0], _p(pre0), None, None, 0, _stream())
            else:
                _call("ctr_tower_gemm_presplit", 0, _p(X), _p(X_lo), _p(tw.P("0.w")), _p(tw.w0_lo),
                      B, tw.sizes[0], Hs[0], _p(acts[0]), _p(tw.P("0.b")),
                      _p(stats[0]) if training else None, 1, _stream())
        else:
            if tw._w0_ready is not None:       # split launched but not needed: still join the fork
                torch.cuda.current_stream().wait_event(tw._w0_ready)
                tw._w0_ready = None
            _call("ctr_tower_layer_fwd", _p(X), tw.sizes[0], tw.sizes[0], None, _p(tw.P("0.w")),
                  _p(tw.P("0.b")), Hs[0], _p(acts[0]), Hs[0], _p(stats[0]) if training else None, 1,
                  B, _stream())
        a = _lib.TowerMidArgs()
        a.L, a.C, a.relu0, a.training = L, len(zs) + 1, 1 if relu0 else 0, 1 if training else 0
        if training:
            dn = [torch.empty((B, H), **f32) for H in Hs]
            dpre = [torch.empty((B, H), **f32) for H in Hs]
            dzs = [torch.empty(B, **f32) for _ in zs]
            dpre0_lo = torch.empty((B, Hs[0]), **f32) if presplit else None
        else:
            dn = dpre = dzs = dpre0_lo = None
        for l, H in enumerate(Hs):
            a.H[l] = H
            if l > 0:
                a.W[l], a.b[l] = _p(tw.P("%d.w" % l)), _p(tw.P("%d.b" % l))
            elif splitk:
                a.b[0], a.pre0 = _p(tw.P("0.b")), _p(pre0)
            if l == 0 and l0 is not None and training:
                a.stats0_part, a.n_stats0_part = _p(l0[1]), l0[1].shape[0]
            a.gamma[l], a.beta[l] = _p(tw.P("%d.bn.gamma" % l)), _p(tw.P("%d.bn.beta" % l))
            a.mean[l], a.var[l] = _p(tw.P("%d.bn.mean" % l)), _p(tw.P("%d.bn.var" % l))
            a.act[l] = _p(acts[l])
            if training:
                a.stats[l] = _p(stats[l])
                a.dgamma[l], a.dbeta[l] = _p(tw.G("%d.bn.gamma" % l)), _p(tw.G("%d.bn.beta" % l))
                a.dbias[l] = _p(tw.G("%d.b" % l))
                a.dn[l], a.dpre[l] = _p(dn[l]), _p(dpre[l])
        a.w_out, a.b_out = _p(tw.P("out.w")), _p(tw.P("out.b"))
        a.state, a.eps, a.p_drop = tw.adam.state_ptr, BN_EPS, tw.dropout if training else 0.0
        a.seed, a.grad_scale = tw.seed & 0xFFFFFFFF, float(grad_scale)
        for c, z in enumerate(zs):
            a.z[c] = _p(z)
            if training:
                a.dz[c] = _p(dzs[c])
        a.hw, a.hb, a.b1 = _p(dense[hw]), _p(dense[hb]), _p(dense[b1]) if relu0 else None
        a.labels, a.y_out, a.logits, a.prob, a.loss = _p(labels), _p(y), _p(logits), _p(prob), _p(loss)
        if training:
            a.dhw, a.dhb = _p(dense[hw].grad), _p(dense[hb].grad)
            a.db1 = _p(dense[b1].grad) if relu0 else None
            a.dw_out, a.db_out = _p(tw.G("out.w")), _p(tw.G("out.b"))
            a.barrier = _p(tw.barrier)
            a.dpre0_lo = _p(dpre0_lo)
        a.timing = _p(tw.timing) if tw.timing is not None else None
        _call("ctr_tower_mid", C.byref(a), B, _stream())
        ctx.tw, ctx.training = tw, training
        tw.last_acts = acts          # post-ReLU hidden activations of the latest call (tests)
        tw.last_y = y                # and the tower's output relu(h . w_out + b_out)
        ctx.saved = (X, acts, stats, dpre, dzs, dn, zs, labels, ws, X_lo, dpre0_lo)
        ctx.l0 = l0
        ctx.mark_non_differentiable(logits, prob)
        return loss, logits, prob

    @staticmethod
    def backward(ctx, gl, _g1, _g2):
        tw = ctx.tw
        if not ctx.training:
            raise RuntimeError("tower_head backward is only defined in training mode")
        X, acts, stats, dpre, dzs, dn, zs, labels, ws, X_lo, dpre0_lo = ctx.saved
        B, dev = X.shape[0], X.device
        L = len(tw.sizes) - 1
        presplit = dpre0_lo is not None
        main, side = torch.cuda.current_stream(), tw.side
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)

        def grad_src(G, H):
            g = _lib.GradSrc()
            g.G, g.ldg, g.a, g.lda, g.kind, g.train, g.eps = _p(G), H, None, 0, 2, 1, BN_EPS
            return g

        fuse_bwd0 = presplit and ctx.l0 is not None and tw.fuse_bwd0

        def dw0_presplit():
            _call("ctr_tower_gemm_presplit", 2, _p(X), _p(X_lo), _p(dpre[0]), _p(dpre0_lo),
                  B, tw.sizes[0], tw.sizes[1], _p(tw.G("0.w")), None, None, 0, side.cuda_stream)

        defer_all = fuse_bwd0 and os.environ.get("CTR_DW_DEFER", "all") == "all"

        def dw_hidden():
            for l in range(L):
                H, K = tw.sizes[l + 1], tw.sizes[l]
                if l == 0 and presplit:
                    if not fuse_bwd0:
                        dw0_presplit()
                    continue
                gs = grad_src(dpre[l], H)
                xin = acts[l - 1] if l > 0 else X
                pro = C.byref(tw.bn_drop(l - 1, stats[l - 1], True)) if l > 0 else None
                _call("ctr_tower_layer_bwd_weights", _p(xin), K, K, pro, C.byref(gs), H,
                      _p(tw.G("%d.w" % l)), None, B, side.cuda_stream)

        if not defer_all:
            with torch.cuda.stream(side):   # dW_l = P(a_{l-1})^T . dpre_l, off the critical path
                dw_hidden()
        H0, K0 = tw.sizes[1], tw.sizes[0]
        gs0 = grad_src(dpre[0], H0)
        if fuse_bwd0:
            # dX = dpre0 . W0^T is formed inside the embedding's scatter kernel (ctr_tower_embed_bwd).
            # Both that kernel and the first layer's weight-gradient GEMM want (nearly) a whole SM's
            # shared memory: the GEMM is queued behind the scatter kernel (``after_scatter``), beside
            # the row optimiser, and only the small hidden-layer kernels run next to the scatter.
            for t_ in dpre + acts + [X, ws, X_lo, dpre0_lo]:
                t_.record_stream(side)

            def after_scatter():
                ev2 = torch.cuda.Event()
                ev2.record(torch.cuda.current_stream())
                side.wait_event(ev2)
                with torch.cuda.stream(side):
                    dw0_presplit()
                    if defer_all:
                        dw_hidden()
                    done = torch.cuda.Event()
                    done.record(side)
                tw._pending = done

            tw.bwd0 = (dpre[0], dpre0_lo, after_scatter)
            done = torch.cuda.Event()
            done.record(side)
            tw._pending = done
            return (tw._zero.expand(B, K0), None, None, None, None, None, None) + tuple(dzs)
        dX = torch.empty((B, K0), dtype=torch.float32, device=dev)
        if presplit:
            _call("ctr_tower_gemm_presplit", 1, _p(dpre[0]), _p(dpre0_lo), _p(tw.P("0.w")),
                  _p(tw.w0_lo), B, K0, H0, _p(dX), None, None, 0, _stream())
        else:
            _call("ctr_tower_layer_bwd_data", C.byref(gs0), H0, _p(tw.P("0.w")), K0, None, None,
                  _p(dX), K0, None, None, B, _stream())
        for t_ in dpre + acts + [X, ws] + ([X_lo, dpre0_lo] if presplit else []):   # read by the side stream
            t_.record_stream(side)
        done = torch.cuda.Event()
        done.record(side)
        tw._pending = done
        return (dX, None, None, None, None, None, None) + tuple(dzs)


class _DcnHeadFn(torch.autograd.Function):
    """ctr_dcn_head: logit = [h | xl] . w + b, prob, mean BCE and - in training - dh, dxl, dw, db in
    the same launch (dcn/dcn.py:151-153,166-169)."""

    @staticmethod
    def forward(ctx, h, xl, anchor, w, dw, hb, dhb, labels, grad_scale, training):
        ctx.set_materialize_grads(False)
        h, xl = h.contiguous(), xl.contiguous()
        B, dev = h.shape[0], h.device
        labels = labels.to(dev, torch.float32).contiguous().view(B)
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        prob = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        dh = torch.empty_like(h) if training else None
        dxl = torch.empty_like(xl) if training else None
        _call("ctr_dcn_head", _p(h), h.shape[1], _p(xl), xl.shape[1], _p(w), _p(hb), _p(labels), B,
              _p(logits), _p(prob), _p(loss), _p(dh), _p(dxl), _p(dw) if training else None,
              _p(dhb) if training else None, float(grad_scale), _stream())
        ctx.g = (dh, dxl)
        ctx.mark_non_differentiable(logits, prob)
        return loss, logits, prob

    @staticmethod
    def backward(ctx, gl, _g1, _g2):
        dh, dxl = ctx.g       # produced with the final scale; the caller backpropagates the loss itself
        return (dh, dxl) + (None,) * 8


def dcn_head(anchor, dense: "DenseParams", h, xl, labels, hw="head.w", hb="head.b", grad_scale=None,
             training=True):
    require_cuda(h, "tower output")
    if grad_scale is None:
        grad_scale = 1.0 / h.shape[0]
    return _DcnHeadFn.apply(h, xl, anchor, dense[hw], dense[hw].grad, dense[hb], dense[hb].grad,
                            labels, grad_scale, training)


def split_lo(x: torch.Tensor) -> torch.Tensor:
    """lo half of the 3xTF32 split of ``x`` (ctr_split_lo)."""
    require_cuda(x, "x")
    x = x.contiguous()
    lo = torch.empty_like(x)
    _call("ctr_split_lo", _p(x), _p(lo), x.numel(), _stream())
    return lo


def tower_head(tw: "FusedTower", X, zs, labels, hw="head.w", hb="head.b", b1="b1", relu0=True,
               grad_scale=None, training=True, X_lo=None):
    """loss, logits, prob of `head([zs..., tower(X)])` (deepfm/deepfm.py:100-129).  ``X_lo``: the
    lo half of X's 3xTF32 split when its producer wrote one (FieldEmbedding.lookup(want_lo=True))."""
    require_cuda(X, "tower input")
    B = X.shape[0]
    if grad_scale is None:
        grad_scale = 1.0 / B
    return _TowerHeadFn.apply(X, tw._anchor, tw, (hw, hb, b1, relu0, grad_scale), labels, training,
                              X_lo, *zs)


class _LossHeadFn(torch.autograd.Function):
    """ctr_loss_head: logits, prob, mean BCE and (training) the input / head-parameter
    gradients in one launch.  The head-parameter gradients are accumulated into the flat
    gradient buffer during the forward; the input gradients are handed to autograd."""

    @staticmethod
    def forward(ctx, anchor, dense, names, relu0, grad_scale, labels, want_grad, *zs):
        ctx.set_materialize_grads(False)
        B, dev = zs[0].shape[0], zs[0].device
        Cn = len(zs)
        zs = [z.contiguous().view(B) for z in zs]
        labels = labels.to(dev, torch.float32).contiguous().view(B)
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        prob = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        zp = (C.c_void_p * Cn)(*[z.data_ptr() for z in zs])
        dzs = [torch.empty(B, dtype=torch.float32, device=dev) for _ in zs] if want_grad else None
        dzp = (C.c_void_p * Cn)(*[d.data_ptr() for d in dzs]) if want_grad else None
        # a head parameter is a name in ``dense`` or an explicit (value, gradient) tensor pair
        (hw, dhw), (hb, dhb), (b1, db1) = [
            n if isinstance(n, tuple) else (dense[n], dense[n].grad) if n in dense else (None, None)
            for n in names]
        _call("ctr_loss_head", zp, dzp, Cn, 1 if relu0 else 0, _p(hw), _p(hb),
              _p(b1) if relu0 else None, _p(labels), B, _p(logits), _p(prob), _p(loss),
              _p(dhw) if want_grad else None, _p(dhb) if want_grad else None,
              _p(db1) if (want_grad and relu0) else None, float(grad_scale), _stream())
        ctx.dzs = dzs
        ctx.mark_non_differentiable(logits, prob)
        return loss, logits, prob

    @staticmethod
    def backward(ctx, gl, _g1, _g2):
        # d(loss)/dz was produced with the final scale (1/(B*world)); the caller backpropagates
        # the loss itself (upstream gradient 1).
        return (None, None, None, None, None, None, None) + tuple(ctx.dzs)


def loss_head(anchor, dense, zs, labels, hw="head.w", hb="head.b", b1="b1", relu0=True,
              grad_scale=None, training=True):
    B = zs[0].shape[0]
    if grad_scale is None:
        grad_scale = 1.0 / B
    return _LossHeadFn.apply(anchor, dense, (hw, hb, b1), relu0, grad_scale, labels, training, *zs)

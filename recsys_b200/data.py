"""Input side of the hot path: the reference's ``input_fn`` contract
(fm/fm.py:100-112: TFRecordDataset -> parse_single_example -> batch ->
[shuffle batches] -> prefetch -> repeat) as a plain Python iterator of
``(features, labels)`` batches, plus a seeded synthetic source of the same shape
for benchmarks.  Batches come out as ``PackedFeatures`` over pinned host buffers
so the model_fn moves them with two async copies.
"""
from __future__ import annotations

import random
import struct
from typing import Iterable, Iterator, List, Sequence

import numpy as np
import torch

from .ops import PackedFeatures


# ------------------------------------------------------------- TFRecord / Example
def iter_tfrecords(path: str) -> Iterator[bytes]:
    """record := u64 len | u32 crc(len) | payload | u32 crc(payload) (crcs not verified)."""
    with open(path, "rb") as f:
        while True:
            hdr = f.read(12)
            if len(hdr) < 12:
                return
            (n,) = struct.unpack("<Q", hdr[:8])
            payload = f.read(n)
            f.read(4)
            yield payload


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if c < 0x80:
            return r, i
        s += 7


def _ld(b, i):
    n, i = _varint(b, i)
    return b[i:i + n], i + n


def parse_example(payload: bytes) -> dict:
    """tf.train.Example -> {key: bytes | float | int | list} (first value of
    single-valued features, list for multi-valued ones)."""
    out = {}
    i = 0
    while i < len(payload):
        tag, i = _varint(payload, i)
        feats, i = _ld(payload, i)
        if tag != 0x0A:
            continue
        j = 0
        while j < len(feats):
            tag, j = _varint(feats, j)
            entry, j = _ld(feats, j)
            if tag != 0x0A:
                continue
            k = 0
            key, feat = None, b""
            while k < len(entry):
                t, k = _varint(entry, k)
                v, k = _ld(entry, k)
                if t == 0x0A:
                    key = v.decode()
                elif t == 0x12:
                    feat = v
            if not feat:
                out[key] = None
                continue
            t, k = _varint(feat, 0)
            lst, k = _ld(feat, k)
            vals: List = []
            m = 0
            if t == 0x0A:       # bytes_list
                while m < len(lst):
                    _, m = _varint(lst, m)
                    v, m = _ld(lst, m)
                    vals.append(bytes(v))
            elif t == 0x12:     # float_list
                while m < len(lst):
                    tt, m = _varint(lst, m)
                    if tt == 0x0A:
                        v, m = _ld(lst, m)
                        vals.extend(struct.unpack("<%df" % (len(v) // 4), v))
                    else:
                        vals.append(struct.unpack("<f", lst[m:m + 4])[0])
                        m += 4
            elif t == 0x1A:     # int64_list
                while m < len(lst):
                    tt, m = _varint(lst, m)
                    if tt == 0x0A:
                        v, m = _ld(lst, m)
                        n = 0
                        while n < len(v):
                            x, n = _varint(v, n)
                            vals.append(x - (1 << 64) if x >= (1 << 63) else x)
                    else:
                        x, m = _varint(lst, m)
                        vals.append(x - (1 << 64) if x >= (1 << 63) else x)
            out[key] = vals
    return out


# --------------------------------------------------------------------- batching
def _pinned(shape, dtype):
    t = torch.empty(shape, dtype=dtype)
    if torch.cuda.is_available():
        t = t.pin_memory()
    return t


def criteo_input_fn(filenames: Sequence[str], batch_size: int, num_epochs: int = -1,
                    need_shuffle: bool = False, shuffle_buffer: int = 1000, id_pipeline=None,
                    seed: int = 0):
    """fm/fm.py:106-112.  Yields (features, labels).  Categorical strings are
    returned raw (numpy object arrays of bytes) unless ``id_pipeline`` is given,
    in which case they are hashed on the device and the batch is PackedFeatures.
    Like the reference, ``shuffle`` acts on whole batches (it follows ``batch``)."""
    from .criteo_schema import cat_feature, cont_feature

    def batches():
        epoch = 0
        while num_epochs < 0 or epoch < num_epochs:
            buf = []
            for fn in filenames:
                for rec in iter_tfrecords(fn):
                    buf.append(parse_example(rec))
                    if len(buf) == batch_size:
                        yield buf
                        buf = []
            if buf:
                yield buf
            epoch += 1

    def to_batch(exs):
        B = len(exs)
        feats = {}
        for k in cont_feature[1:]:
            feats[k] = torch.tensor([e[k][0] for e in exs], dtype=torch.float32).reshape(B, 1)
        for k in cat_feature:
            feats[k] = np.array([e[k][0] if e.get(k) else b"NULL" for e in exs],
                                dtype=object).reshape(B, 1)          # default 'NULL', fm/fm.py:44
        labels = torch.tensor([e["_c0"][0] for e in exs], dtype=torch.float32).reshape(B, 1)
        return feats, labels

    def gen():
        rng = random.Random(seed)
        pool = []
        for exs in batches():
            item = to_batch(exs)
            if not need_shuffle:
                yield item
                continue
            pool.append(item)
            if len(pool) >= shuffle_buffer:
                yield pool.pop(rng.randrange(len(pool)))
        while pool:
            yield pool.pop(rng.randrange(len(pool)))

    return gen()


class SyntheticCriteo:
    """Seeded synthetic 39-field batches: numerics = floor(lognormal),
    ``_c2`` shifted to >= -2; categoricals = pre-hashed local ids, uniform or
    Zipf(1.05) clipped to the bucket count; labels ~ Bernoulli(0.22).  Batches are
    PackedFeatures over pinned host memory (``device=None``) or device tensors."""

    def __init__(self, lay, batch_size: int, n_batches: int, dist: str = "uniform", seed: int = 0,
                 device=None):
        from . import feature_column as fc
        rng = np.random.default_rng(seed)
        self.cont_keys = [c.key for c in lay.columns
                          if isinstance(c.categorical_column, fc.BucketizedColumn)]
        cat_cols = [(c.key, c.num_buckets) for c in lay.columns
                    if not isinstance(c.categorical_column, fc.BucketizedColumn)]
        self.cat_keys = [k for k, _ in cat_cols]
        self.batches = []
        for _ in range(n_batches):
            cont = np.floor(rng.lognormal(1.0, 1.5, size=(batch_size, len(self.cont_keys))))
            for j, k in enumerate(self.cont_keys):
                if k == "_c2":
                    cont[:, j] -= 2.0
            cat = np.empty((batch_size, len(cat_cols)), np.int64)
            for j, (_, n) in enumerate(cat_cols):
                if dist == "zipf":
                    cat[:, j] = np.minimum(rng.zipf(1.05, size=batch_size) - 1, n - 1)
                else:
                    cat[:, j] = rng.integers(0, n, size=batch_size)
            lab = (rng.random((batch_size, 1)) < 0.22).astype(np.float32)
            tc = _pinned(cont.shape, torch.float32).copy_(torch.from_numpy(cont.astype(np.float32)))
            tk = _pinned(cat.shape, torch.int64).copy_(torch.from_numpy(cat))
            tl = _pinned(lab.shape, torch.float32).copy_(torch.from_numpy(lab))
            if device is not None:
                tc, tk, tl = tc.to(device), tk.to(device), tl.to(device)
            self.batches.append((PackedFeatures(tc, tk, self.cont_keys, self.cat_keys), tl))

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)

    def input_fn(self):
        return iter(self.batches)

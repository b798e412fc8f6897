"""Input side of the hot path: the reference's ``input_fn`` contract
(fm/fm.py:100-112: TFRecordDataset -> parse_single_example -> batch ->
[shuffle batches] -> prefetch -> repeat) as a Python iterator of ``(features, labels)``
batches, plus a seeded synthetic source of the same shape for benchmarks.

Framing (masked crc32c verified) and ``tf.train.Example`` decoding run in the library's
multi-threaded host decoder (csrc/records.cu: ctr_tfrecord_scan / ctr_criteo_parse /
ctr_din_parse) straight into pinned batch buffers; the categorical strings are not hashed on the
host - they travel as fixed-width slots and are fingerprinted on the device (ctr_hash_slots) by
the model's id pipeline.  ``iter_tfrecords`` / ``parse_example`` are the plain-Python readers of
the same formats (small files, tests).
"""
from __future__ import annotations

import ctypes as C
import random
import struct
from typing import Iterator, List, Sequence

import numpy as np
import torch

from . import _lib
from .ops import PackedFeatures

CAT_SLOT = 16        # bytes per categorical string slot (Criteo values are 8 hex characters)


# ------------------------------------------------------------- TFRecord / Example
def masked_crc32c(data: bytes) -> int:
    """The checksum TFRecord frames carry: crc32c rotated right by 15 plus 0xa282ead8."""
    a = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, np.uint8)
    return int(_lib.load().ctr_masked_crc32c(a.ctypes.data if len(data) else None, len(data)))


def iter_tfrecords(path: str, verify_crc: bool = True) -> Iterator[bytes]:
    """record := u64 len | u32 crc(len) | payload | u32 crc(payload); crcs verified."""
    with open(path, "rb") as f:
        i = 0
        while True:
            hdr = f.read(12)
            if not hdr:
                return
            if len(hdr) < 12:
                raise ValueError("%s: truncated record header (record %d)" % (path, i))
            (n,) = struct.unpack("<Q", hdr[:8])
            payload = f.read(n)
            tail = f.read(4)
            if len(payload) < n or len(tail) < 4:
                raise ValueError("%s: truncated record %d" % (path, i))
            if verify_crc:
                if masked_crc32c(hdr[:8]) != struct.unpack("<I", hdr[8:])[0] or \
                        masked_crc32c(payload) != struct.unpack("<I", tail)[0]:
                    raise ValueError("%s: crc mismatch in record %d" % (path, i))
            yield payload
            i += 1


class RecordFile:
    """A TFRecord file in memory with its frame index (ctr_tfrecord_scan)."""

    def __init__(self, path: str, verify_crc: bool = True):
        lib = _lib.load()
        self.path = path
        self.buf = np.fromfile(path, dtype=np.uint8)
        n = lib.ctr_tfrecord_scan(self.buf.ctypes.data, self.buf.size, 0, None, None, 0)
        if n < 0:
            raise ValueError("%s: %s" % (path, _lib.last_error()))
        self.off = np.empty(n, np.int64)
        self.len = np.empty(n, np.int32)
        n2 = lib.ctr_tfrecord_scan(self.buf.ctypes.data, self.buf.size, 1 if verify_crc else 0,
                                   self.off.ctypes.data, self.len.ctypes.data, n)
        if n2 < 0:
            raise ValueError("%s: %s" % (path, _lib.last_error()))
        self.n = int(n)

    def __len__(self):
        return self.n


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
                out[key] = []          # an empty Feature: the schema's default applies
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


class CriteoRawBatch(dict):
    """One decoded Criteo batch as the host decoder leaves it: ``cont`` f32 [B,13] (_c1.._c13),
    ``cat_bytes`` u8 [B,26,slot] + ``cat_len`` i32 [B,26] (_c14.._c39, 'NULL' where absent), all
    pinned.  As a features dict it serves the reference's keys: ``_c1.._c13`` -> float [B,1]
    views, ``_c14.._c39`` -> numpy object arrays of bytes [B,1] (materialised on first access;
    the model's id pipeline never needs them - it hashes the slots on the device)."""

    def __init__(self, cont, cat_bytes, cat_len):
        super().__init__()
        self.cont, self.cat_bytes, self.cat_len = cont, cat_bytes, cat_len
        for j in range(13):
            self["_c%d" % (j + 1)] = cont[:, j:j + 1]

    def __missing__(self, key):
        j = int(key[2:]) - 14 if key.startswith("_c") and key[2:].isdigit() else -1
        if not 0 <= j < 26:
            raise KeyError(key)
        b, ln = self.cat_bytes.numpy(), self.cat_len.numpy()
        v = np.array([b[i, j, :ln[i, j]].tobytes() for i in range(b.shape[0])], dtype=object)
        self[key] = v.reshape(-1, 1)
        return self[key]

    def keys(self):
        return ["_c%d" % i for i in range(1, 40)]

    def __contains__(self, key):
        return key in self.keys()

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


def _batches_of(filenames, batch_size, num_epochs, verify_crc):
    """(RecordFile, first record, count) per batch; like ``dataset.batch`` the stream of records
    runs across file boundaries only at a file's end (a short batch closes each epoch)."""
    files = None
    epoch = 0
    while num_epochs < 0 or epoch < num_epochs:
        if files is None:
            files = [RecordFile(fn, verify_crc) for fn in filenames]
        pend = []                                    # [(file, start, count)] of a batch in progress
        need = batch_size
        for rf in files:
            start = 0
            while start < rf.n:
                take = min(need, rf.n - start)
                pend.append((rf, start, take))
                start += take
                need -= take
                if need == 0:
                    yield pend
                    pend, need = [], batch_size
        if pend:
            yield pend
        epoch += 1


def _shuffled(gen, need_shuffle, shuffle_buffer, seed):
    """Like the reference, ``shuffle`` follows ``batch``: it shuffles whole batches."""
    if not need_shuffle:
        yield from gen
        return
    rng = random.Random(seed)
    pool = []
    for item in gen:
        pool.append(item)
        if len(pool) >= shuffle_buffer:
            yield pool.pop(rng.randrange(len(pool)))
    while pool:
        yield pool.pop(rng.randrange(len(pool)))


def criteo_input_fn(filenames: Sequence[str], batch_size: int, num_epochs: int = -1,
                    need_shuffle: bool = False, shuffle_buffer: int = 1000, seed: int = 0,
                    n_threads: int = 0, verify_crc: bool = True):
    """fm/fm.py:106-112.  Yields (features, labels): ``features`` is a ``CriteoRawBatch`` (the
    reference's keys over pinned buffers; categorical strings raw, hashed on the device by the
    model's id pipeline), ``labels`` f32 [B,1] (= _c0)."""
    lib = _lib.load()

    def gen():
        for parts in _batches_of(filenames, batch_size, num_epochs, verify_crc):
            B = sum(c for _, _, c in parts)
            labels, cont = _pinned((B, 1), torch.float32), _pinned((B, 13), torch.float32)
            cat_bytes = _pinned((B, 26, CAT_SLOT), torch.uint8)
            cat_len = _pinned((B, 26), torch.int32)
            done = 0
            for rf, start, cnt in parts:
                rc = lib.ctr_criteo_parse(
                    rf.buf.ctypes.data, rf.off[start:].ctypes.data, rf.len[start:].ctypes.data, cnt,
                    n_threads, labels[done:].data_ptr(), cont[done:].data_ptr(),
                    cat_bytes[done:].data_ptr(), cat_len[done:].data_ptr(), CAT_SLOT)
                if rc != 0:
                    raise ValueError("%s: %s" % (rf.path, _lib.last_error()))
                done += cnt
            yield CriteoRawBatch(cont, cat_bytes, cat_len), labels

    return _shuffled(gen(), need_shuffle, shuffle_buffer, seed)


def din_input_fn(filenames: Sequence[str], batch_size: int, num_epochs: int = -1,
                 need_shuffle: bool = False, shuffle_buffer: int = 1000, seed: int = 0,
                 n_threads: int = 0, verify_crc: bool = True):
    """din/din.py:52-80: label / i_id / i_cate int64 scalars, u_iid_seq / u_icat_seq var-len int64
    densified per record and batched with ``.batch()`` - every record of a batch must carry the
    same history length (a ragged batch raises, as TF's batch op does)."""
    lib = _lib.load()

    def gen():
        for parts in _batches_of(filenames, batch_size, num_epochs, verify_crc):
            B = sum(c for _, _, c in parts)
            rf0, s0, _ = parts[0]
            P = lib.ctr_din_parse(rf0.buf.ctypes.data, rf0.off[s0:].ctypes.data,
                                  rf0.len[s0:].ctypes.data, 1, 1, 0, None, None, None, None, None)
            if P < 0:
                raise ValueError("%s: %s" % (rf0.path, _lib.last_error()))
            labels = _pinned((B,), torch.int64)
            f = {"i_id": _pinned((B,), torch.int64), "i_cate": _pinned((B,), torch.int64),
                 "u_iid_seq": _pinned((B, P), torch.int64), "u_icat_seq": _pinned((B, P), torch.int64)}
            done = 0
            for rf, start, cnt in parts:
                rc = lib.ctr_din_parse(
                    rf.buf.ctypes.data, rf.off[start:].ctypes.data, rf.len[start:].ctypes.data, cnt,
                    n_threads, P, labels[done:].data_ptr(), f["i_id"][done:].data_ptr(),
                    f["i_cate"][done:].data_ptr(), f["u_iid_seq"][done:].data_ptr(),
                    f["u_icat_seq"][done:].data_ptr())
                if rc < 0:
                    raise ValueError("%s: %s" % (rf.path, _lib.last_error()))
                done += cnt
            yield f, labels

    return _shuffled(gen(), need_shuffle, shuffle_buffer, seed)


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

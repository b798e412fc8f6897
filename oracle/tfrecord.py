"""TFRecord framing + ``tf.train.Example`` wire-format reader (oracle).

Test infrastructure - see oracle/__init__.py.  Restates what
``tf.data.TFRecordDataset`` + ``tf.parse_single_example(serial, feature_description)``
do for the reference's ``input_fn`` (fm/fm.py:100-112, xdeepfm/xdeepfm.py:97-120,
din/din.py:52-80) [TF-sem]:
  record  := u64le length | u32le masked_crc32c(length) | payload | u32le masked_crc32c(payload)
  Example := { features(1): Features{ feature(1): map<string, Feature> } }
  Feature := bytes_list(1){value(1)*} | float_list(2){packed f32} | int64_list(3){packed varint}
  FixedLenFeature with a default fills missing keys (``'NULL'`` for _c14.._c39, fm/fm.py:44).
"""
from __future__ import annotations

import struct

import numpy as np

_CRC_TABLE = None


def _crc32c(data: bytes) -> int:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = tbl
    crc = 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = _crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def read_records(path: str, limit: int | None = None, verify_crc: bool = False):
    out = []
    with open(path, "rb") as f:
        while limit is None or len(out) < limit:
            hdr = f.read(12)
            if len(hdr) < 12:
                break
            (n,) = struct.unpack("<Q", hdr[:8])
            payload = f.read(n)
            (crc,) = struct.unpack("<I", f.read(4))
            if verify_crc:
                if struct.unpack("<I", hdr[8:])[0] != masked_crc32c(hdr[:8]):
                    raise ValueError("corrupt TFRecord length crc")
                if crc != masked_crc32c(payload):
                    raise ValueError("corrupt TFRecord payload crc")
            out.append(payload)
    return out


def _varint(buf: bytes, i: int):
    r = 0
    s = 0
    while True:
        b = buf[i]
        i += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, i
        s += 7


def _fields(buf: bytes):
    i = 0
    while i < len(buf):
        key, i = _varint(buf, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = buf[i:i + 8]
            i += 8
        elif wt == 2:
            n, i = _varint(buf, i)
            v = buf[i:i + n]
            i += n
        elif wt == 5:
            v = buf[i:i + 4]
            i += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield fn, wt, v


def parse_example(payload: bytes) -> dict:
    """-> {key: list of bytes | np.float32 array | np.int64 array}."""
    feats = {}
    for fn, _, features in _fields(payload):
        if fn != 1:
            continue
        for fn2, _, entry in _fields(features):
            if fn2 != 1:
                continue
            key, feat = None, b""
            for fn3, _, v in _fields(entry):
                if fn3 == 1:
                    key = v.decode()
                elif fn3 == 2:
                    feat = v
            val = None
            for kind, _, lst in _fields(feat):
                if kind == 1:
                    val = [v for f, _, v in _fields(lst) if f == 1]
                elif kind == 2:
                    vals = []
                    for f, wt, v in _fields(lst):
                        if f == 1 and wt == 2:
                            vals.extend(np.frombuffer(v, "<f4").tolist())
                        elif f == 1:
                            vals.append(struct.unpack("<f", v)[0])
                    val = np.array(vals, np.float32)
                elif kind == 3:
                    vals = []
                    for f, wt, v in _fields(lst):
                        if f == 1 and wt == 2:
                            j = 0
                            while j < len(v):
                                x, j = _varint(v, j)
                                vals.append(x - (1 << 64) if x >= 1 << 63 else x)
                        elif f == 1:
                            vals.append(v - (1 << 64) if v >= 1 << 63 else v)
                    val = np.array(vals, np.int64)
            feats[key] = val
    return feats


def criteo_batch(payloads):
    """parse_single_example with fm/fm.py:43-44's feature_description, batched:
    -> ({_c1.._c13: f32[B,1], _c14.._c39: object(bytes)[B,1]}, labels f32[B,1])."""
    B = len(payloads)
    ex = [parse_example(p) for p in payloads]
    feats = {}
    for i in range(1, 14):
        k = "_c%d" % i
        feats[k] = np.array([e[k][0] for e in ex], np.float32).reshape(B, 1)
    for i in range(14, 40):
        k = "_c%d" % i
        feats[k] = np.array([e[k][0] if e.get(k) else b"NULL" for e in ex],
                            dtype=object).reshape(B, 1)
    labels = np.array([e["_c0"][0] for e in ex], np.float32).reshape(B, 1)
    return feats, labels

"""Criteo 39-field schema and id pipeline of the reference, restated (oracle).

Test infrastructure - see oracle/__init__.py.  Follows
  fm/fm.py:39-44      feature_description (13 float + 26 string, default 'NULL')
  fm/fm.py:47-97      build_feature_columns (= xdeepfm/xdeepfm.py:44-94, dcn/dcn.py:49-99)
and the TF feature-column semantics those lines invoke [TF-sem]:
  * ``numeric_column(normalizer_fn=log(x+1))`` (``log(x+4)`` for ``_c2``,
    fm/fm.py:76-78) -> ``bucketized_column(boundaries)``: id = number of
    boundaries <= value (``Bucketize`` == upper_bound), fp32 arithmetic;
  * ``categorical_column_with_hash_bucket(key, n)``: id = Fingerprint64(bytes) mod n;
  * ``input_layer`` concatenates its columns sorted by column *name*.
"""
from __future__ import annotations

import numpy as np

from .farmhash import hash_bucket_fast

CONT = ["_c%d" % i for i in range(1, 14)]          # fm/fm.py:39 minus the label _c0 (:48)
CAT = ["_c%d" % i for i in range(14, 40)]          # fm/fm.py:40

# fm/fm.py:54-67 - quantile boundaries (raw scale, applied to the log value).
BOUNDARIES = [
    [0.0, 1.0, 2.0, 3.0, 5.0, 12.0],
    [0.0, 1.0, 2.0, 4.0, 10.0, 28.0, 76.0, 301.0],
    [1.0, 2.0, 3.0, 5.0, 7.0, 10.0, 16.0, 24.0, 54.0],
    [1.0, 2.0, 3.0, 5.0, 6.0, 9.0, 13.0, 20.0],
    [20.0, 155.0, 1087.0, 1612.0, 2936.0, 5064.0, 8622.0, 16966.0, 39157.0],
    [3.0, 7.0, 13.0, 24.0, 36.0, 53.0, 85.0, 154.0, 411.0],
    [0.0, 1.0, 2.0, 4.0, 6.0, 10.0, 17.0, 43.0],
    [1.0, 2.0, 4.0, 6.0, 8.0, 12.0, 17.0, 25.0, 37.0],
    [4.0, 8.0, 16.0, 28.0, 41.0, 63.0, 109.0, 147.0, 321.0],
    [0.0, 1.0, 2.0],
    [0.0, 1.0, 2.0, 3.0, 4.0, 8.0],
    [0.0, 1.0, 2.0],
    [1.0, 2.0, 3.0, 5.0, 7.0, 10.0, 14.0, 22.0],
]
LOG_OFFSET = [4.0 if n == "_c2" else 1.0 for n in CONT]        # fm/fm.py:76-78

# fm/fm.py:72-73 - the effective (second) assignment; "R-ref".
HASH_BUCKETS = [1460, 583, 100000, 100000, 305, 23, 12517, 633, 3, 93145, 5683, 100000, 3194, 27,
                14992, 100000, 10, 5652, 2172, 3, 100000, 17, 15, 100000, 104, 100000]
# fm/fm.py:69-70 - the overwritten first assignment = true cardinalities; "R-full".
HASH_BUCKETS_FULL = [1460, 583, 10131226, 2202607, 305, 23, 12517, 633, 3, 93145, 5683, 8351592,
                     3194, 27, 14992, 5461305, 10, 5652, 2172, 3, 7046546, 17, 15, 286180, 104,
                     142571]


def column_name(key: str, kind: str = "embedding") -> str:
    """TF column names: ``_c1_bucketized_embedding`` / ``_c14_embedding``
    (``_indicator`` for the linear one-hots; bare key for a numeric column)."""
    if kind == "numeric":
        return key
    return (key + "_bucketized_" + kind) if key in CONT else (key + "_" + kind)


class CriteoSpec:
    """Field order, rows per field and row offsets of the concatenated table.

    ``fields`` is the order of the F axis everywhere (kernel, E[B,F*D], w1):
    ``input_layer``'s sorted-column-name order [TF-sem], i.e.
    _c10,_c11,_c12,_c13,_c14.._c19,_c1,_c20.._c29,_c2,_c30.._c39,_c3.._c9.
    """

    def __init__(self, full_cardinality: bool = False):
        hb = HASH_BUCKETS_FULL if full_cardinality else HASH_BUCKETS
        rows = {k: len(b) + 1 for k, b in zip(CONT, BOUNDARIES)}
        rows.update({k: n for k, n in zip(CAT, hb)})
        self.fields = sorted(CONT + CAT, key=column_name)
        self.rows = [rows[k] for k in self.fields]
        self.offsets = np.concatenate([[0], np.cumsum(self.rows)]).astype(np.int64)
        self.total_rows = int(self.offsets[-1])
        self.is_cont = [k in CONT for k in self.fields]
        self.cont_fields = [k for k in self.fields if k in CONT]   # F-axis order, numeric only
        self.F = len(self.fields)

    def boundaries(self, key):
        return BOUNDARIES[CONT.index(key)]

    def log_offset(self, key):
        return LOG_OFFSET[CONT.index(key)]


def log_normalise(x: np.ndarray, key: str) -> np.ndarray:
    """fm/fm.py:76-78: tf.log(x + 1.0), tf.log(x + 4.0) for _c2; fp32."""
    off = np.float32(LOG_OFFSET[CONT.index(key)])
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.log(np.asarray(x, np.float32) + off).astype(np.float32)


def bucketize(v: np.ndarray, boundaries) -> np.ndarray:
    """TF ``Bucketize``: index = #boundaries <= v (std::upper_bound). NaN -> len(b)
    (upper_bound's comparator ``v < b`` is false for NaN on every element)."""
    b = np.asarray(boundaries, np.float32)
    v = np.asarray(v, np.float32)
    out = np.searchsorted(b, v, side="right").astype(np.int64)
    return np.where(np.isnan(v), len(b), out)


def criteo_rows(features: dict, spec: CriteoSpec) -> np.ndarray:
    """features -> global row ids int64 [B, F] in ``spec.fields`` order.

    ``features[_c1.._c13]``: float arrays [B] or [B,1];
    ``features[_c14.._c39]``: sequences of bytes (raw strings, hashed here),
    or integer arrays of already-hashed local ids.
    """
    cols = []
    for f, key in enumerate(spec.fields):
        val = features[key]
        if key in CONT:
            x = np.asarray(val, np.float32).reshape(-1)
            local = bucketize(log_normalise(x, key), spec.boundaries(key))
        else:
            arr = np.asarray(val)
            if arr.dtype.kind in "iu":
                local = arr.reshape(-1).astype(np.int64)
            else:
                flat = np.asarray(val, dtype=object).reshape(-1)
                local = np.array([hash_bucket_fast(bytes(s), spec.rows[f]) for s in flat], np.int64)
        cols.append(local + spec.offsets[f])
    return np.stack(cols, axis=1)


def criteo_logx(features: dict, spec: CriteoSpec) -> np.ndarray:
    """[B, 13] fp32 log-normalised numerics in ``spec.cont_fields`` order: the
    13 ``numeric_column`` inputs of xdeepfm's linear part (xdeepfm/xdeepfm.py:82)."""
    return np.stack([log_normalise(np.asarray(features[k], np.float32).reshape(-1), k)
                     for k in spec.cont_fields], axis=1)


def synthetic_features(B: int, seed: int = 0, spec: CriteoSpec | None = None,
                       dist: str = "zipf", hashed: bool = True):
    """Seeded synthetic Criteo-shaped batch.

    Numerics: rounded log-normal counts (``_c2`` shifted to >= -2 as in the shard);
    categoricals: local ids ~ Zipf(1.05) clipped (or uniform) when ``hashed``,
    else 8-hex-char byte strings (10 % missing -> b'NULL', fm/fm.py:44);
    labels ~ Bernoulli(0.22).
    """
    spec = spec or CriteoSpec()
    rng = np.random.default_rng(seed)
    feats = {}
    for k in CONT:
        x = np.floor(rng.lognormal(1.0, 1.5, size=B)).astype(np.float32)
        if k == "_c2":
            x = x - 2.0
        feats[k] = x.reshape(B, 1)
    for k in CAT:
        n = spec.rows[spec.fields.index(k)]
        if hashed:
            if dist == "zipf":
                ids = np.minimum(rng.zipf(1.05, size=B) - 1, n - 1)
            else:
                ids = rng.integers(0, n, size=B)
            feats[k] = ids.astype(np.int64).reshape(B, 1)
        else:
            raw = rng.integers(0, 1 << 32, size=B, dtype=np.uint64)
            strs = [b"%08x" % int(v) for v in raw]
            miss = rng.random(B) < 0.1
            feats[k] = np.array([b"NULL" if m else s for s, m in zip(strs, miss)],
                                dtype=object).reshape(B, 1)
    labels = (rng.random(B) < 0.22).astype(np.float32).reshape(B, 1)
    return feats, labels

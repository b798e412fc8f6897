"""Lightweight mirror of the ``tf.feature_column`` objects the reference builds in
``build_feature_columns`` (fm/fm.py:47-97, xdeepfm/xdeepfm.py:44-94,
dcn/dcn.py:49-99, deepfm/deepfm.py:37-51).  The objects only carry the facts the
id pipeline and the table layout need (name, key, boundaries, bucket count,
dimension); the arithmetic happens in libctr_b200 (ctr_criteo_rows).

``layout(columns)`` resolves a list of embedding columns into the F-axis order
TF's ``input_layer`` uses - columns sorted by ``name`` - with per-field row
counts and offsets into the single concatenated table.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence


@dataclass(frozen=True)
class NumericColumn:
    key: str
    log_offset: Optional[float] = None      # normalizer_fn = log(x + log_offset)
    normalizer_fn: Optional[Callable] = field(default=None, compare=False)

    @property
    def name(self):
        return self.key


@dataclass(frozen=True)
class BucketizedColumn:
    source_column: NumericColumn
    boundaries: tuple

    @property
    def key(self):
        return self.source_column.key

    @property
    def name(self):
        return self.source_column.key + "_bucketized"

    @property
    def num_buckets(self):
        return len(self.boundaries) + 1


@dataclass(frozen=True)
class HashedCategoricalColumn:
    key: str
    hash_bucket_size: int
    dtype: str = "string"

    @property
    def name(self):
        return self.key

    @property
    def num_buckets(self):
        return self.hash_bucket_size


@dataclass(frozen=True)
class EmbeddingColumn:
    categorical_column: object
    dimension: int
    combiner: str = "mean"

    @property
    def key(self):
        return self.categorical_column.key

    @property
    def name(self):
        return self.categorical_column.name + "_embedding"

    @property
    def num_buckets(self):
        return self.categorical_column.num_buckets


@dataclass(frozen=True)
class IndicatorColumn:
    categorical_column: object

    @property
    def key(self):
        return self.categorical_column.key

    @property
    def name(self):
        return self.categorical_column.name + "_indicator"

    @property
    def num_buckets(self):
        return self.categorical_column.num_buckets


def _probe_log_offset(fn) -> float:
    """Recover ``off`` from a reference-style ``lambda x: log(x + off)`` by probing
    it with plain floats through a tiny shim exposing ``log``."""
    try:
        v = fn(_Probe(0.0))
        return float(v.offset)
    except Exception as e:  # pragma: no cover - defensive
        raise ValueError("normalizer_fn must be of the form log(x + c)") from e


class _Probe:
    def __init__(self, off):
        self.offset = off

    def __add__(self, c):
        return _Probe(self.offset + float(c))

    __radd__ = __add__


def log(x):
    """Stand-in for ``tf.log`` inside a reference-style normalizer lambda."""
    if isinstance(x, _Probe):
        return x
    return math.log(x)


def numeric_column(key, normalizer_fn=None, log_offset=None):
    if log_offset is None and normalizer_fn is not None:
        log_offset = _probe_log_offset(normalizer_fn)
    return NumericColumn(key, log_offset, normalizer_fn)


def bucketized_column(source_column, boundaries):
    b = tuple(float(x) for x in boundaries)
    if list(b) != sorted(b):
        raise ValueError("boundaries must be sorted")
    return BucketizedColumn(source_column, b)


def categorical_column_with_hash_bucket(key, hash_bucket_size, dtype="string"):
    if hash_bucket_size < 1:
        raise ValueError("hash_bucket_size must be at least 1")
    return HashedCategoricalColumn(key, int(hash_bucket_size), dtype)


def embedding_column(categorical_column, dimension, combiner="mean"):
    return EmbeddingColumn(categorical_column, int(dimension), combiner)


def indicator_column(categorical_column):
    return IndicatorColumn(categorical_column)


@dataclass
class Layout:
    """F-axis resolution of a list of embedding columns."""
    columns: List[EmbeddingColumn]          # sorted by name (input_layer order)
    keys: List[str]
    rows: List[int]
    offsets: List[int]                      # len F+1
    dimension: int

    @property
    def F(self):
        return len(self.columns)

    @property
    def total_rows(self):
        return self.offsets[-1]

    def field_of(self, key):
        return self.keys.index(key)


def layout(embedding_columns: Sequence[EmbeddingColumn]) -> Layout:
    cols = sorted(embedding_columns, key=lambda c: c.name)
    dims = {c.dimension for c in cols}
    if len(dims) != 1:
        raise ValueError("all embedding columns must share one dimension, got %s" % sorted(dims))
    rows = [c.num_buckets for c in cols]
    offs = [0]
    for r in rows:
        offs.append(offs[-1] + r)
    return Layout(cols, [c.key for c in cols], rows, offs, dims.pop())


def first_order_fields(linear_columns, lay: Layout):
    """(bitmask of F-axis fields that have an indicator column, list of numeric keys)
    for a reference-style ``linear_feature_columns`` list."""
    mask = 0
    numeric = []
    for c in linear_columns:
        if isinstance(c, IndicatorColumn):
            mask |= 1 << lay.field_of(c.key)
        elif isinstance(c, NumericColumn):
            numeric.append(c.key)
        else:
            raise ValueError("unsupported linear column %r" % (c,))
    return mask, numeric

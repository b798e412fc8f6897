"""Seeded synthetic inputs for the DIN path (oracle side):
history length ~ Uniform{1..P}, zero padded (id 0 = padding, din/din.py:107), ids
Zipf-ish over 1..n-1, labels Bernoulli(0.3)."""
from __future__ import annotations

import numpy as np


def synthetic_din(B: int, P: int = 100, seed: int = 0, n_items: int = 63002, n_cates: int = 802,
                  zipf: bool = True):
    rng = np.random.default_rng(seed)

    def ids(n, size):
        if zipf:
            return (np.minimum(rng.zipf(1.2, size=size), n - 1)).astype(np.int64)
        return rng.integers(1, n, size=size).astype(np.int64)

    lens = rng.integers(1, P + 1, size=B)
    mask = np.arange(P)[None, :] < lens[:, None]
    feats = {
        "i_id": ids(n_items, B),
        "i_cate": ids(n_cates, B),
        "u_iid_seq": ids(n_items, (B, P)) * mask,
        "u_icat_seq": ids(n_cates, (B, P)) * mask,
    }
    labels = (rng.random(B) < 0.3).astype(np.int64)
    return feats, labels

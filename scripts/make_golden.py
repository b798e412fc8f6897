"""Generate the committed fixtures under tests/golden/.

Run in the authoring container (needs /root/reference for the real shard):
    python scripts/make_golden.py

  criteo_shard256.tfrecord  the first 256 TFRecord frames of the shard, byte for byte
  criteo_shard256.npz   first 256 records of the reference's only data fixture,
                        xdeepfm/part-r-00000, parsed by oracle/tfrecord.py (crc-verified):
                        13 numerics, 26 raw byte strings ('NULL' default), labels - plus the
                        row ids oracle/criteo.py derives from them.
  oracle_<model>.npz    seeded oracle (fp64) known answers per model: logits, loss and
                        per-parameter gradient checksums.  The reference has no golden
                        vectors (parity unpinned): these freeze the oracle against
                        regressions, they are not an external truth.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import criteo, models, synth, tfrecord  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SHARD = "/root/reference/xdeepfm/part-r-00000"

# small-table Criteo-shaped spec used by the oracle known answers (keeps dense grads tiny)
SMALL_HASH = [min(n, 997) for n in criteo.HASH_BUCKETS]


def small_spec():
    spec = criteo.CriteoSpec()
    rows = {k: len(b) + 1 for k, b in zip(criteo.CONT, criteo.BOUNDARIES)}
    rows.update({k: n for k, n in zip(criteo.CAT, SMALL_HASH)})
    spec.rows = [rows[k] for k in spec.fields]
    spec.offsets = np.concatenate([[0], np.cumsum(spec.rows)]).astype(np.int64)
    spec.total_rows = int(spec.offsets[-1])
    return spec


def criteo_batch(B, seed, spec):
    feats, labels = criteo.synthetic_features(B, seed=seed, spec=spec, dist="zipf")
    rows = criteo.criteo_rows(feats, spec)
    logx = criteo.criteo_logx(feats, spec)
    return feats, labels, rows, logx


def model_batch(model, B, seed, spec):
    feats, labels, rows, logx = criteo_batch(B, seed, spec)
    batch = {"rows": torch.from_numpy(rows), "labels": torch.from_numpy(labels)}
    if model == "xdeepfm":
        batch["logx"] = torch.from_numpy(logx.astype(np.float64))
        batch["cat_mask"] = torch.tensor([0.0 if c else 1.0 for c in spec.is_cont],
                                         dtype=torch.float64)
    return feats, batch


def summarise(out, grads):
    d = {"logits": out["logits"].numpy(), "loss": np.array(float(out["loss"]))}
    for k, g in grads.items():
        g = g.numpy()
        d["gsum." + k] = np.array(g.sum())
        d["gabs." + k] = np.array(np.abs(g).sum())
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    if os.path.exists(SHARD):
        recs = tfrecord.read_records(SHARD, limit=256, verify_crc=True)
        feats, labels = tfrecord.criteo_batch(recs)
        spec = criteo.CriteoSpec()
        rows = criteo.criteo_rows(feats, spec)
        save = {"labels": labels, "rows": rows}
        for k in criteo.CONT:
            save[k] = feats[k]
        for k in criteo.CAT:
            save[k] = np.array([bytes(v) for v in feats[k].reshape(-1)], dtype="S8")
        np.savez_compressed(os.path.join(OUT, "criteo_shard256.npz"), **save)
        print("criteo_shard256.npz", rows.shape, float(labels.mean()))
        # the same 256 records as raw TFRecord frames (byte copy of the head of the shard), so the
        # product's decoder (recsys_b200/data.py, csrc/records.cu) can be tested where
        # /root/reference is absent
        import struct
        with open(SHARD, "rb") as f:
            blob = bytearray()
            for _ in range(256):
                hdr = f.read(12)
                (n,) = struct.unpack("<Q", hdr[:8])
                blob += hdr + f.read(n + 4)
        with open(os.path.join(OUT, "criteo_shard256.tfrecord"), "wb") as f:
            f.write(bytes(blob))
        print("criteo_shard256.tfrecord", len(blob), "bytes")
    spec = small_spec()
    for model in ("fm", "deepfm", "xdeepfm", "dcn"):
        kw = {}
        if model == "xdeepfm":
            kw = dict(cin_layers=(16, 8))
        p = models.init_params(model, spec.total_rows, deep_layers=(32, 16), seed=3, **kw)
        _, batch = model_batch(model, 64, 7, spec)
        out, grads = models.loss_and_grads(model, p, batch)
        np.savez_compressed(os.path.join(OUT, "oracle_%s.npz" % model), **summarise(out, grads))
        print(model, float(out["loss"]))
    feats, labels = synth.synthetic_din(32, P=20, seed=5, n_items=500, n_cates=50)
    p = models.init_params("din", D=16, seed=3, din_items=500, din_cates=50)
    batch = {k: torch.from_numpy(v) for k, v in feats.items()}
    batch["labels"] = torch.from_numpy(labels)
    out, grads = models.loss_and_grads("din", p, batch)
    np.savez_compressed(os.path.join(OUT, "oracle_din.npz"), **summarise(out, grads))
    print("din", float(out["loss"]))


if __name__ == "__main__":
    main()

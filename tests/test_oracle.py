"""CPU: the oracle against its pins - the TF-documented hash example, the real
TFRecord shard frozen under tests/golden/, the committed known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import criteo, farmhash, models, synth, tfsem

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_farmhash_tf_documented_example():
    # tf.strings.to_hash_bucket_fast(["Hello", "TensorFlow", "2.x"], 3) -> [0, 2, 2]  (TF API docs)
    got = [farmhash.hash_bucket_fast(s, 3) for s in (b"Hello", b"TensorFlow", b"2.x")]
    assert got == [0, 2, 2]


def test_farmhash_length_branches_are_deterministic_and_distinct():
    seen = set()
    for n in (0, 1, 3, 4, 7, 8, 16, 17, 32, 33, 64, 65, 128, 200):
        s = bytes((i * 7 + 3) % 251 for i in range(n))
        h = farmhash.fingerprint64(s)
        assert 0 <= h < 1 << 64
        assert h == farmhash.fingerprint64(bytes(s))
        seen.add(h)
    assert len(seen) == 14


def test_field_order_is_sorted_column_name_order():
    spec = criteo.CriteoSpec()
    assert spec.fields[:5] == ["_c10", "_c11", "_c12", "_c13", "_c14"]
    assert spec.fields[10] == "_c1" and spec.fields[21] == "_c2" and spec.fields[-1] == "_c9"
    assert spec.total_rows == 840646                         # 108 + 840538
    assert criteo.CriteoSpec(full_cardinality=True).total_rows == 108 + 33762565


def test_bucketize_upper_bound_semantics():
    b = [0.0, 1.0, 2.0]
    v = np.array([-1.0, 0.0, 0.5, 1.0, 2.0, 5.0, np.nan], np.float32)
    assert criteo.bucketize(v, b).tolist() == [0, 1, 1, 2, 3, 3, 3]
    # _c2 uses log(x+4): x=-2 -> log 2 = .693 -> bucket 1 of [0,1,2,...]
    assert criteo.bucketize(criteo.log_normalise(np.array([-2.0]), "_c2"),
                            criteo.BOUNDARIES[1]).tolist() == [1]


def test_real_shard_fixture_rows():
    z = np.load(os.path.join(GOLD, "criteo_shard256.npz"))
    spec = criteo.CriteoSpec()
    feats = {k: z[k] for k in criteo.CONT}
    feats.update({k: np.array([bytes(v) for v in z[k]], dtype=object) for k in criteo.CAT})
    rows = criteo.criteo_rows(feats, spec)
    assert np.array_equal(rows, z["rows"])
    assert abs(float(z["labels"].mean()) - 0.21875) < 1e-9
    f5 = spec.fields.index("_c5")        # every sample lands in bucket 0 of _c5
    assert (rows[:, f5] == spec.offsets[f5]).all()
    assert ((rows >= spec.offsets[:-1]) & (rows < spec.offsets[1:])).all()


def test_auc_matches_exact_roc_auc():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    y = (rng.random(5000) < 0.3)
    p = np.clip(rng.normal(0.3 + 0.2 * y, 0.15), 0, 1).astype(np.float32)
    m = tfsem.StreamingAUC()
    for i in range(0, 5000, 1000):
        m.update(y[i:i + 1000], p[i:i + 1000])
    assert abs(m.result() - roc_auc_score(y, p)) < 2e-3
    acc = tfsem.StreamingAccuracy()
    acc.update(y, p)
    assert abs(acc.result() - float(((p > 0.5) == y).mean())) < 1e-3


def test_tf_adam_first_step_and_dense_decay():
    p = {"w": torch.tensor([[1.0, 2.0], [3.0, 4.0]], dtype=torch.float64)}
    opt = tfsem.TFAdam(p, lr=0.1)
    g = {"w": torch.tensor([[0.5, -0.5], [0.0, 0.0]], dtype=torch.float64)}
    opt.step(p, g)
    # first step: m/(sqrt(v)+eps) with bias correction folded into lr_t -> |step| = lr (eps tiny)
    assert torch.allclose(p["w"][0], torch.tensor([0.9, 2.1], dtype=torch.float64), atol=1e-6)
    assert torch.equal(p["w"][1], torch.tensor([3.0, 4.0], dtype=torch.float64))
    opt.step(p, {"w": torch.zeros(2, 2, dtype=torch.float64)})
    assert p["w"][0, 0] < 0.9            # momentum keeps moving rows with zero gradient [TF-sem]


@pytest.mark.parametrize("model", ["fm", "deepfm", "xdeepfm", "dcn", "din"])
def test_known_answers(model):
    import make_golden as mg
    z = np.load(os.path.join(GOLD, "oracle_%s.npz" % model))
    if model == "din":
        feats, labels = synth.synthetic_din(32, P=20, seed=5, n_items=500, n_cates=50)
        p = models.init_params("din", D=16, seed=3, din_items=500, din_cates=50)
        batch = {k: torch.from_numpy(v) for k, v in feats.items()}
        batch["labels"] = torch.from_numpy(labels)
    else:
        spec = mg.small_spec()
        kw = dict(cin_layers=(16, 8)) if model == "xdeepfm" else {}
        p = models.init_params(model, spec.total_rows, deep_layers=(32, 16), seed=3, **kw)
        _, batch = mg.model_batch(model, 64, 7, spec)
    out, grads = models.loss_and_grads(model, p, batch)
    assert np.allclose(out["logits"].numpy(), z["logits"], rtol=1e-10, atol=1e-12)
    assert abs(float(out["loss"]) - float(z["loss"])) < 1e-12
    for k, g in grads.items():
        assert abs(float(g.sum()) - float(z["gsum." + k])) <= 1e-9 * (1 + abs(float(z["gabs." + k])))


def test_cin_literal_matches_einsum():
    g = torch.Generator().manual_seed(0)
    B, m, D = 5, 7, 4
    X0 = torch.randn(B, m, D, generator=g, dtype=torch.float64)
    p = {"cin.0.w": torch.randn(m * m, 6, generator=g, dtype=torch.float64),
         "cin.0.b": torch.randn(6, generator=g, dtype=torch.float64),
         "cin.1.w": torch.randn(m * 6, 3, generator=g, dtype=torch.float64),
         "cin.1.b": torch.zeros(3, dtype=torch.float64),
         "cin.out.w": torch.randn(9, 1, generator=g, dtype=torch.float64),
         "cin.out.b": torch.zeros(1, dtype=torch.float64)}
    y = models.cin(p, X0)
    x1 = torch.relu(torch.einsum("bid,bjd,ijh->bhd", X0, X0, p["cin.0.w"].view(m, m, 6))
                    + p["cin.0.b"][None, :, None])
    x2 = torch.relu(torch.einsum("bid,bjd,ijh->bhd", X0, x1, p["cin.1.w"].view(m, 6, 3)))
    ref = torch.relu(torch.cat([x1, x2], 1).sum(-1) @ p["cin.out.w"])
    assert torch.allclose(y, ref, atol=1e-12)


def test_fp32_oracle_close_to_fp64():
    import make_golden as mg
    spec = mg.small_spec()
    p = models.init_params("deepfm", spec.total_rows, deep_layers=(32, 16), seed=3)
    _, batch = mg.model_batch("deepfm", 64, 7, spec)
    o64 = models.deepfm(p, **batch, training=True)
    p32 = {k: v.float() for k, v in p.items()}
    o32 = models.deepfm(p32, **batch, training=True)
    assert torch.allclose(o32["logits"].double(), o64["logits"], rtol=1e-4, atol=1e-5)

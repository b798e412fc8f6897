"""GPU parity, model_fn level: the five reference entry points against the fp64 oracle on
the same seeded Criteo-/DIN-shaped inputs.  Tolerance from BASELINE north_star: logits
<= 1e-4 relative (plus 1e-5 absolute for values near zero); loss to 1e-5; gradients to
1e-3 of their max-norm; AUC / logloss equal to 4 decimals."""
import os

import numpy as np
import pytest
import torch

import make_golden as mg
from oracle import criteo, models as om, synth, tfsem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _features_to_torch(feats):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in feats.items()}


def _build(model, spec, cuda, deep_layers="32,16", cin_layers="16,8", **extra):
    from recsys_b200 import _core
    from recsys_b200 import criteo_schema as cs
    from recsys_b200.estimator import VariableStore
    hb = [spec.rows[spec.fields.index(k)] for k in criteo.CAT]
    linear = {"fm": "indicator_all", "deepfm": "indicator_all", "xdeepfm": "numeric+indicator",
              "dcn": "numeric"}[model]
    lin, emb = cs.build_columns(16, linear=linear, hash_buckets=hb)
    params = {"linear_feature_columns": lin, "embedding_feature_columns": emb,
              "embedding_size": 16, "learning_rate": 1e-3, "dropout": 0.0,
              "deep_layers": deep_layers, "cross_layers": cin_layers if model == "xdeepfm" else 4,
              "variable_store": VariableStore(), "device": cuda}
    params.update(extra)
    cls = {"fm": _core.FMModel, "deepfm": _core.DeepFMModel, "xdeepfm": _core.XDeepFMModel,
           "dcn": _core.DCNModel}[model]
    m = params["variable_store"].get(model, lambda: cls(params))
    return m, params


def _model_fn(model):
    import importlib
    return importlib.import_module("recsys_b200.%s.%s" % (model, model)).model_fn


def _check_grads(m, grads64, tol=1e-3):
    bad = []

    def close(a, b, what):
        err = float((a.detach().cpu().double() - b).abs().max())
        s = float(b.abs().max()) + 1e-12
        if not err <= tol * s + 1e-7:
            bad.append("%s: err %.3e scale %.3e" % (what, err, s))
    close(m.emb.dtable, grads64["emb"], "d emb")
    if m.emb.with_w1 and "w1" in grads64:
        close(m.emb.dw1, grads64["w1"], "d w1")
    dg = m.dense_grads()
    for k, g in grads64.items():
        if k in ("emb", "w1", "emb_dnn"):
            continue
        if k.startswith("cross."):
            l, wb = k.split(".")[1:]
            close(dg["cross." + wb][int(l)], g, k)
        else:
            close(dg[k], g.reshape(dg[k].shape), k)
    assert not bad, "; ".join(bad)


@pytest.mark.parametrize("model", ["fm", "deepfm", "dcn", "xdeepfm", "xdeepfm-fp32", "fm-torch",
                                   "deepfm-torch", "xdeepfm-torch", "dcn-torch"])
@pytest.mark.parametrize("B", [64, 1000])
def test_criteo_models_match_oracle(cuda, model, B):
    """Default = fused tower / loss-head kernels; ``-torch`` = the torch (cuBLAS) tower."""
    prec = "fp32" if model.endswith("-fp32") else "tf32x3"
    fused = not model.endswith("-torch")
    model = model.split("-")[0]
    spec = mg.small_spec()
    kw = dict(cin_layers=(16, 8)) if model == "xdeepfm" else {}
    p64 = om.init_params(model, spec.total_rows, deep_layers=(32, 16), seed=3, **kw)
    feats, batch = mg.model_batch(model, B, 7, spec)
    out64, g64 = om.loss_and_grads(model, p64, batch)
    m, params = _build(model, spec, cuda, cin_precision=prec, fused_tower=fused)
    m.load_state(p64)
    spec_ = _model_fn(model)(_features_to_torch(feats), batch["labels"], "train", params)
    logits = m.last["logits"].detach().cpu().double().reshape(-1)
    ref = out64["logits"].reshape(-1)
    assert float(((logits - ref).abs() / (ref.abs() + 0.1)).max()) <= 1e-4
    assert abs(float(spec_.loss) - float(out64["loss"])) <= 1e-5
    assert spec_.predictions["prob"].shape == out64["prob"].shape        # [B,1] or [B] as the reference
    m.backward(m.last["loss"])
    _check_grads(m, g64)
    if model == "xdeepfm":
        def close(a, b):
            return float((a.cpu().double() - b).abs().max()) <= 1e-3 * float(b.abs().max()) + 1e-7
        assert close(m.emb_dnn.dtable, g64["emb_dnn"])


def _dropout_masks(m, B):
    """The keep masks the fused tower will use at the current step: run the BN+dropout prologue
    on an all-ones activation with identity BN (mean 0, var 1-eps, gamma 1, beta 0)."""
    import ctypes as C
    from recsys_b200 import _lib
    lib = _lib.load()
    masks = []
    tw = m.tower
    for l, H in enumerate(tw.sizes[1:]):
        ones = torch.ones(B, H, device=m.device)
        out = torch.empty_like(ones)
        zero, one = torch.zeros(H, device=m.device), torch.ones(H, device=m.device)
        d = _lib.BnDrop()
        d.sums, d.mean, d.var = None, zero.data_ptr(), (one - 1e-3).data_ptr()
        d.gamma, d.beta, d.state = one.data_ptr(), zero.data_ptr(), tw.adam.state_ptr
        d.eps, d.p_drop, d.seed, d.layer, d.enabled = 1e-3, tw.dropout, tw.seed, tw.layer_base + l, 1
        var = one - 1e-3
        d.var = var.data_ptr()
        rc = lib.ctr_bn_drop_apply(ones.data_ptr(), H, C.byref(d), out.data_ptr(), B,
                                   torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        keep = (out > 0).float().cpu()
        assert torch.allclose(out.cpu(), keep / (1 - tw.dropout), atol=1e-5)
        masks.append(keep)
    return masks


@pytest.mark.parametrize("model", ["deepfm", "dcn"])
def test_fused_tower_dropout_matches_oracle_with_same_masks(cuda, model):
    """dropout 0.5 (the reference's training default): the fused kernels regenerate the mask in
    the backward from (seed, layer, step); with those masks injected into the oracle, logits,
    loss and every gradient agree; the keep rate is ~0.5 and masks differ between layers/steps.
    DCN: the tower ends in BN + dropout (no final dense layer, dcn/dcn.py:144-149)."""
    spec = mg.small_spec()
    B = 512
    p64 = om.init_params(model, spec.total_rows, deep_layers=(32, 16), seed=3)
    feats, batch = mg.model_batch(model, B, 9, spec)
    m, params = _build(model, spec, cuda, dropout=0.5, fused_tower=True)
    m.load_state(p64)
    masks = _dropout_masks(m, B)
    assert all(0.4 < float(k.mean()) < 0.6 for k in masks) and not torch.equal(masks[0][:, :16], masks[1])
    out64, g64 = om.loss_and_grads(model, p64, batch, dropout=0.5, masks=[k.double() for k in masks])
    sp = _model_fn(model)(_features_to_torch(feats), batch["labels"], "train", params)
    logits = m.last["logits"].detach().cpu().double().reshape(-1)
    ref = out64["logits"].reshape(-1)
    assert float(((logits - ref).abs() / (ref.abs() + 0.1)).max()) <= 1e-4
    assert abs(float(sp.loss) - float(out64["loss"])) <= 1e-5
    m.backward(m.last["loss"])
    _check_grads(m, g64)
    m.apply_gradients()                               # advances the device step counter
    masks2 = _dropout_masks(m, B)
    assert not torch.equal(masks[0], masks2[0])


def test_modes_and_estimator_spec_contract(cuda):
    """fm/fm.py:135-170: PREDICT -> predictions+export_outputs, EVAL -> +loss+metrics,
    TRAIN -> +train_op; eval-mode BN uses the (never updated) moving stats."""
    from recsys_b200.deepfm import deepfm
    from recsys_b200.estimator import ModeKeys
    spec = mg.small_spec()
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(32, 16), seed=3)
    feats, batch = mg.model_batch("deepfm", 200, 11, spec)
    m, params = _build("deepfm", spec, cuda)
    m.load_state(p64)
    tf = _features_to_torch(feats)
    sp = deepfm.model_fn(tf, None, ModeKeys.PREDICT, params)
    assert sp.loss is None and sp.train_op is None and "serving_default" in sp.export_outputs
    se = deepfm.model_fn(tf, batch["labels"], ModeKeys.EVAL, params)
    o = om.deepfm(p64, **batch, training=False)
    assert torch.allclose(se.predictions["prob"].cpu().double(), o["prob"], rtol=1e-4, atol=1e-6)
    assert abs(float(se.loss) - float(o["loss"])) <= 1e-5
    auc, acc = tfsem.StreamingAUC(), tfsem.StreamingAccuracy()
    auc.update(batch["labels"].numpy(), o["prob"].numpy())
    acc.update(batch["labels"].numpy(), o["prob"].numpy())
    assert round(se.eval_metric_ops["AUC"].result(), 4) == round(auc.result(), 4)
    assert round(se.eval_metric_ops["Accuracy"].result(), 4) == round(acc.result(), 4)
    st = deepfm.model_fn(tf, batch["labels"], ModeKeys.TRAIN, params)
    assert callable(st.train_op) and params["variable_store"].global_step == 0
    st.train_op()
    assert params["variable_store"].global_step == 1


@pytest.mark.parametrize("mode", ["exact_tf", "lazy", "lazy-fused"])
def test_training_steps_follow_tf_adam(cuda, mode):
    """Three optimiser steps of DeepFM == oracle forward/backward + tfsem.TFAdam
    (dense-decay semantics for exact_tf; touched rows only for lazy; lazy-fused = scatter-add and
    row Adam in one pass, ctr_embed_bwd_adam)."""
    from recsys_b200.deepfm import deepfm
    spec = mg.small_spec()
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(32, 16), seed=3)
    fused = mode.endswith("-fused")
    mode = mode.split("-")[0]
    m, params = _build("deepfm", spec, cuda, embedding_adam=mode, learning_rate=1e-2,
                       fused_row_adam=fused)
    m.load_state(p64)
    train = {k: v for k, v in p64.items() if not k.endswith((".bn.mean", ".bn.var"))}
    opt = tfsem.TFAdam(train, lr=1e-2)
    for step in range(3):
        feats, batch = mg.model_batch("deepfm", 128, 20 + step, spec)
        out, g = om.loss_and_grads("deepfm", p64, batch)
        lazy = {"emb": batch["rows"].reshape(-1), "w1": batch["rows"].reshape(-1)} \
            if mode == "lazy" else None
        opt.step(train, g, lazy_rows=lazy)
        p64.update(train)
        sp = deepfm.model_fn(_features_to_torch(feats), batch["labels"], "train", params)
        assert abs(float(sp.loss) - float(out["loss"])) <= 2e-4 * (1 + step)
        sp.train_op()
    assert torch.allclose(m.emb.table.cpu().double(), p64["emb"], rtol=1e-3, atol=2e-4)
    assert torch.allclose(m.dense["dnn.0.w"].detach().cpu().double(), p64["dnn.0.w"], rtol=1e-3,
                          atol=2e-4)


def test_real_shard_fm_logloss_and_auc(cuda):
    """BASELINE config 1 on the reference's own records (raw strings hashed on the device):
    logloss and AUC equal to 4 decimals between the CUDA path and the oracle."""
    from recsys_b200 import _core
    from recsys_b200.estimator import VariableStore
    from recsys_b200.fm import fm
    z = np.load(os.path.join(GOLD, "criteo_shard256.npz"))
    spec = criteo.CriteoSpec()
    p64 = om.init_params("fm", spec.total_rows, seed=1)
    rows = torch.from_numpy(z["rows"])
    labels = torch.from_numpy(z["labels"])
    o = om.fm(p64, rows, labels)
    lin, emb = fm.build_feature_columns(16)
    params = {"linear_feature_columns": lin, "embedding_feature_columns": emb, "embedding_size": 16,
              "learning_rate": 1e-3, "dropout": 0.5, "variable_store": VariableStore(), "device": cuda}
    m = params["variable_store"].get("fm", lambda: _core.FMModel(params))
    m.load_state(p64)
    feats = {k: torch.from_numpy(z[k]) for k in criteo.CONT}
    feats.update({k: np.array([bytes(v) for v in z[k]], dtype=object).reshape(-1, 1)
                  for k in criteo.CAT})
    se = fm.model_fn(feats, labels, "eval", params)
    assert round(float(se.loss), 4) == round(float(o["loss"]), 4)
    auc = tfsem.StreamingAUC()
    auc.update(labels.numpy(), o["prob"].numpy())
    assert round(se.eval_metric_ops["AUC"].result(), 4) == round(auc.result(), 4)
    assert se.predictions["prob"].shape == (256, 1)


@pytest.mark.parametrize("B,P", [(32, 20), (300, 100)])
def test_din_matches_oracle(cuda, B, P):
    from recsys_b200 import _core
    from recsys_b200.din import din
    from recsys_b200.estimator import VariableStore
    feats, labels = synth.synthetic_din(B, P=P, seed=5, n_items=500, n_cates=50)
    p64 = om.init_params("din", D=16, seed=3, din_items=500, din_cates=50)
    g = torch.Generator().manual_seed(1)
    p64["i_item"] = torch.randn(500, generator=g, dtype=torch.float64) * 0.1
    for k in list(p64):
        if k.endswith(".b"):
            p64[k] = torch.randn(p64[k].shape, generator=g, dtype=torch.float64) * 0.05
    batch = {k: torch.from_numpy(v) for k, v in feats.items()}
    batch["labels"] = torch.from_numpy(labels)
    out64, g64 = om.loss_and_grads("din", p64, batch)
    params = {"embedding_size": 16, "learning_rate": 1e-3, "dropout": 0.0, "din_items": 500,
              "din_cates": 50, "variable_store": VariableStore(), "device": cuda}
    m = params["variable_store"].get("din", lambda: _core.DINModel(params))
    m.load_state(p64)
    tfeat = {k: torch.from_numpy(v) for k, v in feats.items()}
    sp = din.model_fn(tfeat, torch.from_numpy(labels), "train", params)
    logits = m.last["logits"].detach().cpu().double()
    ref = out64["logits"]
    assert float(((logits - ref).abs() / (ref.abs() + 0.1)).max()) <= 1e-4
    assert abs(float(sp.loss) - float(out64["loss"])) <= 1e-5
    m.backward(m.last["loss"])

    def close(a, b, what, tol=1e-3):
        err = float((a.detach().cpu().double() - b).abs().max())
        s = float(b.abs().max()) + 1e-12
        assert err <= tol * s + 1e-7, "%s: err %.3e scale %.3e" % (what, err, s)
    close(m.emb.dtable[:500], g64["i_id"], "d i_id")
    close(m.emb.dtable[500:], g64["i_cate"], "d i_cate")
    close(m.emb.dw1[:500], g64["i_item"], "d i_item")
    dg = m.dense_grads()
    for k, gg in g64.items():
        if k in dg:
            close(dg[k], gg.reshape(dg[k].shape), k)


def test_din_dropout_matches_oracle_with_same_masks(cuda):
    """din/din.py:15,118,136: the reference's default dropout 0.5 sits inside both activation
    units (after the 80- and the 40-wide layer, per history position) and after every MLP layer.
    The kernels draw counter-based masks (Philox) and regenerate them in the backward; with the
    same masks injected into the oracle, logits, loss and every gradient agree."""
    from recsys_b200 import _core, ops
    from recsys_b200.din import din
    from recsys_b200.estimator import VariableStore
    B, P = 192, 40
    feats, labels = synth.synthetic_din(B, P=P, seed=6, n_items=500, n_cates=50)
    p64 = om.init_params("din", D=16, seed=4, din_items=500, din_cates=50)
    g = torch.Generator().manual_seed(2)
    p64["i_item"] = torch.randn(500, generator=g, dtype=torch.float64) * 0.1
    for k in list(p64):
        if k.endswith(".b"):
            p64[k] = torch.randn(p64[k].shape, generator=g, dtype=torch.float64) * 0.05
    params = {"embedding_size": 16, "learning_rate": 1e-3, "dropout": 0.5, "din_items": 500,
              "din_cates": 50, "variable_store": VariableStore(), "device": cuda}
    m = params["variable_store"].get("din", lambda: _core.DINModel(params))
    m.load_state(p64)
    masks = {}
    for unit, name in enumerate(("att_iid", "att_cat")):
        scales = ops.din_dropout_masks(m.att_opts(unit, True), B * P, cuda)
        masks[name] = [(s > 0).double().cpu() for s in scales]
        assert all(torch.allclose(s[s > 0], torch.full_like(s[s > 0], 2.0)) for s in scales)
        assert all(0.45 < float(k.mean()) < 0.55 for k in masks[name])
    assert not torch.equal(masks["att_iid"][0], masks["att_cat"][0])
    masks["mlp"] = [k.double() for k in _dropout_masks(m, B)]
    assert all(0.4 < float(k.mean()) < 0.6 for k in masks["mlp"])
    batch = {k: torch.from_numpy(v) for k, v in feats.items()}
    batch["labels"] = torch.from_numpy(labels)
    out64, g64 = om.loss_and_grads("din", p64, batch, dropout=0.5, masks=masks)
    sp = din.model_fn({k: torch.from_numpy(v) for k, v in feats.items()}, torch.from_numpy(labels),
                      "train", params)
    logits = m.last["logits"].detach().cpu().double()
    ref = out64["logits"]
    assert float(((logits - ref).abs() / (ref.abs() + 0.1)).max()) <= 1e-4
    assert abs(float(sp.loss) - float(out64["loss"])) <= 1e-5
    m.backward(m.last["loss"])

    def close(a, b, what, tol=1e-3):
        err = float((a.detach().cpu().double() - b).abs().max())
        s = float(b.abs().max()) + 1e-12
        assert err <= tol * s + 1e-7, "%s: err %.3e scale %.3e" % (what, err, s)
    close(m.emb.dtable[:500], g64["i_id"], "d i_id")
    close(m.emb.dtable[500:], g64["i_cate"], "d i_cate")
    dg = m.dense_grads()
    for k, gg in g64.items():
        if k in dg:
            close(dg[k], gg.reshape(dg[k].shape), k)
    # eval mode: no dropout, deterministic
    e1 = din.model_fn({k: torch.from_numpy(v) for k, v in feats.items()}, torch.from_numpy(labels),
                      "eval", params)
    o = om.din(p64, **batch, training=False)
    assert torch.allclose(e1.predictions["prob"].cpu().double(), o["prob"], rtol=1e-4, atol=1e-6)


def test_din_out_of_range_ids_are_flagged_not_read(cuda):
    """ids >= table rows must not read / RED out of bounds.  History ids outside the
    table are treated as padding, target ids are wrapped, and the sticky status word makes the
    next eval call raise."""
    from recsys_b200 import _core
    from recsys_b200.din import din
    from recsys_b200.estimator import VariableStore
    B, P = 64, 12
    feats, labels = synth.synthetic_din(B, P=P, seed=7, n_items=500, n_cates=50)
    params = {"embedding_size": 16, "learning_rate": 1e-3, "dropout": 0.0, "din_items": 500,
              "din_cates": 50, "variable_store": VariableStore(), "device": cuda}
    tf = {k: torch.from_numpy(v.copy()) for k, v in feats.items()}
    clean = din.model_fn(tf, torch.from_numpy(labels), "eval", params).predictions["prob"].clone()
    m = params["variable_store"]._objs["din"]
    bad = {k: v.clone() for k, v in tf.items()}
    was_pad = bad["u_iid_seq"][:, 3] == 0
    bad["u_iid_seq"][:, 3] = torch.where(was_pad, torch.tensor(10 ** 7), bad["u_iid_seq"][:, 3])
    assert bool(was_pad.any())
    with torch.no_grad():        # the out-of-range ids replaced padding and are read as padding
        _, prob, _ = m.forward(bad, torch.from_numpy(labels), False)
    assert torch.equal(clean, prob)
    with pytest.raises(RuntimeError, match="history id outside"):
        m.check_status()
    assert int(m.ids.status.item()) == 0          # reading the word clears it
    bad2 = {k: v.clone() for k, v in tf.items()}
    bad2["i_id"][0] = 500
    with pytest.raises(RuntimeError, match="categorical id outside"):
        din.model_fn(bad2, torch.from_numpy(labels), "eval", params)


def test_deepfm_two_field_int64_keys_are_hashed(cuda):
    """deepfm/deepfm.py:37-51: u_id / i_id are int64 keys; categorical_column_with_hash_bucket
    hashes their decimal strings (as_string + Fingerprint64) [TF-sem].  The ids the model looks
    up must be hash(str(key)) % buckets, not key % buckets."""
    from recsys_b200 import _core
    from recsys_b200.deepfm import deepfm
    from recsys_b200.estimator import VariableStore
    from oracle import farmhash
    lin, emb = deepfm.build_model_columns(16)
    params = {"linear_feature_columns": lin, "embedding_feature_columns": emb, "embedding_size": 16,
              "learning_rate": 1e-3, "dropout": 0.0, "deep_layers": "32,16",
              "variable_store": VariableStore(), "device": cuda}
    rng = np.random.default_rng(0)
    u = rng.integers(-5, 10 ** 12, size=(300, 1))
    i = rng.integers(0, 10 ** 6, size=(300, 1))
    u[0, 0], u[1, 0] = 0, -17
    feats = {"u_id": torch.from_numpy(u), "i_id": torch.from_numpy(i)}
    sp = deepfm.model_fn(feats, torch.zeros(300), "eval", params)
    m = params["variable_store"]._objs["deepfm"]
    order = [c.key for c in m.lay.columns]
    want = {"u_id": [farmhash.fingerprint64(str(int(v)).encode()) % 500000 for v in u[:, 0]],
            "i_id": [farmhash.fingerprint64(str(int(v)).encode()) % 100000 for v in i[:, 0]]}
    rows = m.rows.cpu().numpy()
    for f, k in enumerate(order):
        assert (rows[:, f] - m.lay.offsets[f] == np.array(want[k])).all(), k
    assert sp.predictions["prob"].shape == (300,)


def test_graphed_step_with_host_batches_follows_the_oracle(cuda):
    """estimator.GraphedTrainStep fed from pinned host batches (the bench's e2e path: H2D on a
    copy stream overlapped with the previous replay, one D2D into the static buffers, loss read
    back on a third stream): the per-step losses follow the oracle's TF-Adam trajectory."""
    from recsys_b200.deepfm import deepfm
    from recsys_b200.estimator import GraphedTrainStep
    spec = mg.small_spec()
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(32, 16), seed=3)
    train = {k: v for k, v in p64.items() if not k.endswith((".bn.mean", ".bn.var"))}
    opt = tfsem.TFAdam(train, lr=1e-2)
    nsteps, B = 6, 256
    data = [mg.model_batch("deepfm", B, 40 + s, spec) for s in range(nsteps)]
    host = [({k: v.pin_memory() for k, v in _features_to_torch(f).items()},
             torch.as_tensor(b["labels"]).pin_memory()) for f, b in data]
    side = torch.cuda.Stream(device=cuda)
    with torch.cuda.stream(side):
        m, params = _build("deepfm", spec, cuda, learning_rate=1e-2)
        m.load_state(p64)
        step = GraphedTrainStep(deepfm.model_fn, params, host[0][0], host[0][1], warmup=1)
        # the warm-up ran one real training step: put parameters and optimiser state back
        assert m.emb.record
        D = m.emb.D
        m.emb.rec[:, D:4 * D].zero_()           # m | v | g
        m.emb.rec[:, 4 * D + 1:].zero_()        # m1 v1 g1 | claim
        m.dense.m.zero_()
        m.dense.v.zero_()
        m.dense.grad.zero_()
        m.adam.reset()
        m.load_state(p64)
        losses = torch.zeros(nsteps).pin_memory()
        for s in range(nsteps):
            step(*host[s])
            step.loss_to_host(losses[s:s + 1].view(()))
    torch.cuda.synchronize()
    for s in range(nsteps):
        out, g = om.loss_and_grads("deepfm", p64, data[s][1])
        assert abs(float(losses[s]) - float(out["loss"])) <= 3e-4 * (1 + s), (s, float(losses[s]))
        opt.step(train, g, lazy_rows={"emb": data[s][1]["rows"].reshape(-1),
                                      "w1": data[s][1]["rows"].reshape(-1)})
        p64.update(train)


def test_graphed_step_pipeline_equals_eager_steps(cuda):
    """The pipelined step - two graphs over two input buffers, the batch as ONE pinned blob
    (pin_batch / run_device_batch), ids computed on the copy stream (PackedFeatures.rows), lookup
    fused with the first tower layer, row and dense optimiser kernels side by side - against
    the same model stepped eagerly through model_fn on the same batches: identical loss trajectory
    and identical tables / dense weights at the end (same kernels, same arithmetic; dropout on)."""
    from recsys_b200 import feature_column as fc, ops
    from recsys_b200.data import SyntheticCriteo
    from recsys_b200.deepfm import deepfm
    from recsys_b200.estimator import GraphedTrainStep
    spec = mg.small_spec()
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(32, 16), seed=3)
    nsteps, B = 7, 512
    results = []
    for mode in ("eager", "graphs"):
        side = torch.cuda.Stream(device=cuda)
        with torch.cuda.stream(side):
            m, params = _build("deepfm", spec, cuda, learning_rate=1e-3, dropout=0.5)
            m.load_state(p64)
            lay = fc.layout(params["embedding_feature_columns"])
            host = SyntheticCriteo(lay, B, nsteps, dist="zipf", seed=11, device=None).batches
            losses = torch.zeros(nsteps).pin_memory()
            if mode == "eager":
                for s in range(nsteps):
                    f, l = host[s]
                    pf = ops.PackedFeatures(f.cont.to(cuda), f.cat.to(cuda), f.cont_keys, f.cat_keys)
                    sp = deepfm.model_fn(pf, l.to(cuda), "train", params)
                    sp.train_op()
                    losses[s] = float(sp.loss)
            else:
                step = GraphedTrainStep(deepfm.model_fn, params, host[0][0], host[0][1], warmup=1)
                assert step.nbuf == 2 and step._prefetch is not None
                D = m.emb.D                      # undo the warm-up step (see the test above)
                m.emb.rec[:, D:4 * D].zero_()
                m.emb.rec[:, 4 * D + 1:].zero_()
                m.dense.m.zero_()
                m.dense.v.zero_()
                m.dense.grad.zero_()
                m.adam.reset()
                m.load_state(p64)
                blobs = [step.pin_batch(f, l) for f, l in host]
                for s in range(nsteps):
                    step.wait_loss_slot()
                    step.run_device_batch(blobs[s])
                    step.loss_to_host(losses[s:s + 1].view(()))
        torch.cuda.synchronize()
        results.append((losses.clone(), m.emb.table.clone(), m.dense.flat.clone(),
                        int(m.adam.state.view(torch.int32)[0])))
    (le, te, de, ne), (lg, tg, dg, ng) = results
    assert ne == nsteps and ng == nsteps           # the device schedule advanced once per step
    assert torch.allclose(le, lg, rtol=1e-4, atol=1e-6), (le, lg)
    # parameters: equal up to the summation order of the atomics; Adam turns a sign flip of a
    # noise-level gradient element into a full lr-sized difference, so a handful of elements may
    # differ by a few lr - everything else agrees to 1e-5
    for a, b in ((te, tg), (de, dg)):
        diff = (a - b).abs()
        assert float((diff > 1e-5).float().mean()) <= 2e-3
        assert float(diff.max()) <= 2e-2

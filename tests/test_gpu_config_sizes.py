"""GPU parity at the BASELINE.json configuration sizes (not toy sizes): the model_fn entry points
against the fp64 oracle on the same seeded inputs.

  config 2  DeepFM   B=4096, deep_layers '100,100', the reference's capped table (840 646 rows)
  config 3  xDeepFM  B=8192, cross_layers '128,128', deep_layers '100,100', CIN in 3xTF32 (tcgen05)
  config 4  DIN      B=4096, history length 100, tables 63 002 / 802 rows

Tolerance (BASELINE north_star): logits <= 1e-4 relative (+1e-5 absolute near zero, written as
|d| / (|ref| + 0.1)); loss to 1e-5; gradients to 1e-3 of their max-norm.  Parity is against the
oracle's restatement of the reference graphs - the reference itself (TF 1.x) cannot run here, so
the oracle is unpinned."""
import os

import numpy as np
import pytest
import torch

import make_golden as mg
from oracle import criteo, models as om, synth, tfsem as T

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    a = a.detach().cpu().double().reshape(-1)
    ref = ref.reshape(-1)
    return float(((a - ref).abs() / (ref.abs() + 0.1)).max())


def _features_to_torch(feats):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in feats.items()}


def _criteo_params(model, cuda, **extra):
    import importlib
    from recsys_b200.estimator import VariableStore
    mod = importlib.import_module("recsys_b200.%s.%s" % (model, model))
    lin, emb = mod.build_feature_columns(16)               # the reference's own column lists
    params = {"linear_feature_columns": lin, "embedding_feature_columns": emb, "embedding_size": 16,
              "learning_rate": 1e-3, "dropout": 0.0, "deep_layers": "100,100",
              "variable_store": VariableStore(), "device": cuda}
    params.update(extra)
    return mod, params


def _close(a, b, what, tol=1e-3):
    err = float((a.detach().cpu().double() - b).abs().max())
    s = float(b.abs().max()) + 1e-12
    assert err <= tol * s + 1e-7, "%s: err %.3e scale %.3e" % (what, err, s)


def test_deepfm_config2_matches_oracle(cuda):
    """BASELINE configs[1]: DeepFM Criteo 39-field emb16 batch 4096, MLP 100,100, fwd + bwd."""
    from recsys_b200 import _core
    spec = criteo.CriteoSpec()
    B = 4096
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(100, 100), seed=11)
    feats, batch = mg.model_batch("deepfm", B, 21, spec)
    out64, g64 = om.loss_and_grads("deepfm", p64, batch)
    mod, params = _criteo_params("deepfm", cuda)
    m = params["variable_store"].get("deepfm", lambda: _core.DeepFMModel(params))
    m.load_state(p64)
    assert m.tower.use_mid and m.emb.R == spec.total_rows == 840646
    sp = mod.model_fn(_features_to_torch(feats), batch["labels"], "train", params)
    assert _rel(m.last["logits"], out64["logits"]) <= 1e-4
    assert abs(float(sp.loss) - float(out64["loss"])) <= 1e-5
    m.backward(m.last["loss"])
    _close(m.emb.dtable, g64["emb"], "d emb")
    _close(m.emb.dw1, g64["w1"], "d w1")
    dg = m.dense_grads()
    for k, g in g64.items():
        if k in dg:
            _close(dg[k], g.reshape(dg[k].shape), k)


def _xdeepfm_oracle_logits(p, batch, chunk=512):
    """oracle.models.xdeepfm (xdeepfm/xdeepfm.py:123-233), with the CIN evaluated in sample chunks:
    its literal op order materialises [D, B, 39, Hp] (5 GB in fp64 at this size); CIN has no
    cross-sample term, the BN of the tower (batch statistics) is evaluated on the whole batch."""
    rows, logx, cat_mask = batch["rows"], batch["logx"], batch["cat_mask"]
    with torch.no_grad():
        E = om.gather(p["emb"], rows)
        lin = (p["w1"][rows] * cat_mask).sum(1, keepdim=True) + logx @ p["wnum"].reshape(-1, 1)
        linear_y = torch.relu(lin + p["b1"])
        cin_y = torch.cat([om.cin(p, E[i:i + chunk]) for i in range(0, E.shape[0], chunk)])
        Ed = om.gather(p["emb_dnn"], rows)
        h = om.dnn_tower(p, Ed.reshape(Ed.shape[0], -1), 2, True, 0.0, None)
        dnn_y = T.dense(h, p["dnn.out.w"], p["dnn.out.b"], relu=True)
        logits = T.dense(torch.cat([linear_y, cin_y, dnn_y], -1), p["head.w"], p["head.b"])
        z = batch["labels"].to(logits.dtype).reshape(logits.shape)
        loss = T.sigmoid_cross_entropy_with_logits(logits, z).mean()
    return logits, loss


@pytest.mark.parametrize("prec,tol", [("tf32x3", 1e-4), ("tf32", None)])
def test_xdeepfm_config3_matches_oracle(cuda, prec, tol):
    """BASELINE configs[2]: xDeepFM CIN=[128,128] emb16 batch 8192, the reference's duplicated
    tables.  tf32x3 is the parity mode (logits <= 1e-4 rel); plain tf32 is the labelled
    lower-precision extra: its measured end-to-end logit error is printed and only sanity-bounded."""
    from recsys_b200 import _core
    spec = criteo.CriteoSpec()
    B = 8192
    p64 = om.init_params("xdeepfm", spec.total_rows, deep_layers=(100, 100), cin_layers=(128, 128),
                         seed=12)
    feats, batch = mg.model_batch("xdeepfm", B, 22, spec)
    ref_logits, ref_loss = _xdeepfm_oracle_logits(p64, batch)
    mod, params = _criteo_params("xdeepfm", cuda, cross_layers="128,128", cin_precision=prec)
    m = params["variable_store"].get("xdeepfm", lambda: _core.XDeepFMModel(params))
    m.load_state(p64)
    sp = mod.model_fn(_features_to_torch(feats), batch["labels"], "train", params)
    rel = _rel(m.last["logits"], ref_logits)
    msg = "xdeepfm config3 %s: max logit rel err %.3e, loss err %.3e" \
        % (prec, rel, abs(float(sp.loss) - float(ref_loss)))
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):      # keep the measured figure (pytest hides the stdout of passing tests)
        with open(os.path.join(out, "xdeepfm_config3_logit_err.txt"), "a") as f:
            f.write(msg + "\n")
    assert sp.predictions["prob"].shape == (B, 1)
    if tol is not None:
        assert rel <= tol
        assert abs(float(sp.loss) - float(ref_loss)) <= 1e-5
    else:
        assert rel <= 5e-2            # plain tf32 does not meet 1e-4; see the printed figure
    m.backward(m.last["loss"])        # the backward runs at this size (values: kernel-level tests)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(m.dense.grad).all())


def test_din_config4_matches_oracle(cuda):
    """BASELINE configs[3]: DIN seq_len 100, emb16, batch 4096, tables 63 002 / 802 rows."""
    from recsys_b200 import _core
    from recsys_b200.din import din
    from recsys_b200.estimator import VariableStore
    B, P = 4096, 100
    feats, labels = synth.synthetic_din(B, P=P, seed=15)
    p64 = om.init_params("din", D=16, seed=13)
    g = torch.Generator().manual_seed(1)
    p64["i_item"] = torch.randn(p64["i_item"].shape, generator=g, dtype=torch.float64) * 0.1
    for k in list(p64):
        if k.endswith(".b"):
            p64[k] = torch.randn(p64[k].shape, generator=g, dtype=torch.float64) * 0.05
    batch = {k: torch.from_numpy(v) for k, v in feats.items()}
    batch["labels"] = torch.from_numpy(labels)
    out64, g64 = om.loss_and_grads("din", p64, batch)
    params = {"embedding_size": 16, "learning_rate": 1e-3, "dropout": 0.0,
              "variable_store": VariableStore(), "device": cuda}
    m = params["variable_store"].get("din", lambda: _core.DINModel(params))
    assert (m.n_items, m.n_cates) == (63002, 802)
    m.load_state(p64)
    sp = din.model_fn({k: torch.from_numpy(v) for k, v in feats.items()}, torch.from_numpy(labels),
                      "train", params)
    assert _rel(m.last["logits"], out64["logits"]) <= 1e-4
    assert abs(float(sp.loss) - float(out64["loss"])) <= 1e-5
    m.backward(m.last["loss"])
    _close(m.emb.dtable[:63002], g64["i_id"], "d i_id")
    _close(m.emb.dtable[63002:], g64["i_cate"], "d i_cate")
    _close(m.emb.dw1[:63002], g64["i_item"], "d i_item")
    dg = m.dense_grads()
    for k, gg in g64.items():
        if k in dg:
            _close(dg[k], gg.reshape(dg[k].shape), k)

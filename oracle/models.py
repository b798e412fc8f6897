"""The five reference ``model_fn`` graphs restated op for op on torch-CPU (oracle).

Test infrastructure - see oracle/__init__.py.  dtype float64 = ground truth,
float32 = the timed "reference CPU path".  Gradients come from torch autograd
over these forwards, exactly as the reference gets them from
``optimizer.minimize`` (fm/fm.py:162-163).

Inputs are already-resolved global row ids ``rows[B,F]`` (oracle/criteo.py turns
raw features into them).  Parameters are a flat ``{name: tensor}`` dict; the
names are shared by convention with recsys_b200 (DESIGN.md "parameter names"):

  emb [R,D]  w1 [R]  b1 [1]                      tables (F-axis order), first-order
  dnn.{l}.w [in,out] dnn.{l}.b  dnn.{l}.bn.{gamma,beta,mean,var}   tower
  dnn.out.w [last,1] dnn.out.b  head.w [k,1] head.b
  xdeepfm: emb_dnn (second input_layer's tables), wnum [13], cin.{k}.w [39*Hp,H], cin.{k}.b,
           cin.out.w [sumH,1], cin.out.b
  dcn:     cross.{l}.w [FD], cross.{l}.b [FD]
  din:     i_item [63002], i_id [63002,E], i_cate [802,E],
           att_iid.{0,1,2}.{w,b}, att_cat.{0,1,2}.{w,b}, mlp.{0,1,2}.{w,b}, mlp.out.{w,b}
"""
from __future__ import annotations

import torch

from . import tfsem as T

DIN_ITEMS, DIN_CATES = 63002, 802        # din/din.py:88-90 (hard-coded)
DIN_ATT_LAYERS = [80, 40]                # din/din.py:85
DIN_MLP_LAYERS = [100, 50, 20]           # din/din.py:86


# ------------------------------------------------------------------ parameters
def _tower(p, prefix, sizes, gen, dtype, bn):
    for l, (i, o) in enumerate(zip(sizes[:-1], sizes[1:])):
        p[f"{prefix}.{l}.w"] = T.glorot_uniform((i, o), gen, dtype)
        p[f"{prefix}.{l}.b"] = torch.zeros(o, dtype=dtype)
        if bn:
            p[f"{prefix}.{l}.bn.gamma"] = torch.ones(o, dtype=dtype)
            p[f"{prefix}.{l}.bn.beta"] = torch.zeros(o, dtype=dtype)
            p[f"{prefix}.{l}.bn.mean"] = torch.zeros(o, dtype=dtype)
            p[f"{prefix}.{l}.bn.var"] = torch.ones(o, dtype=dtype)


def init_params(model: str, total_rows: int = 0, F: int = 39, D: int = 16,
                deep_layers=(100, 100), cin_layers=(128, 128), cross_layers: int = 4,
                share_embeddings: bool = False, seed: int = 0, dtype=torch.float64,
                din_items: int = DIN_ITEMS, din_cates: int = DIN_CATES):
    """Variables each model_fn creates, initialised as TF would [TF-sem].
    Non-zero biases are drawn so that parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    if model in ("fm", "deepfm", "xdeepfm", "dcn"):
        p["emb"] = T.truncated_normal((total_rows, D), D ** -0.5, g, dtype)
    if model in ("fm", "deepfm", "xdeepfm"):
        # dense(linear one-hots, 1): kernel [R,1] glorot-uniform, bias [1] (fm/fm.py:121)
        p["w1"] = T.glorot_uniform((total_rows, 1), g, dtype).reshape(-1) * 50.0
        p["b1"] = torch.full((1,), 0.05, dtype=dtype)
    if model == "fm":
        p["head.w"] = T.glorot_uniform((2, 1), g, dtype)
        p["head.b"] = torch.zeros(1, dtype=dtype)
    if model in ("deepfm", "xdeepfm"):
        _tower(p, "dnn", [F * D] + list(deep_layers), g, dtype, bn=True)
        p["dnn.out.w"] = T.glorot_uniform((deep_layers[-1], 1), g, dtype)
        p["dnn.out.b"] = torch.full((1,), 0.1, dtype=dtype)
        p["head.w"] = T.glorot_uniform((3, 1), g, dtype)
        p["head.b"] = torch.zeros(1, dtype=dtype)
    if model == "xdeepfm":
        if not share_embeddings:
            p["emb_dnn"] = T.truncated_normal((total_rows, D), D ** -0.5, g, dtype)
        p["wnum"] = T.glorot_uniform((13, 1), g, dtype).reshape(-1)
        hp = F
        for k, h in enumerate(cin_layers):
            p[f"cin.{k}.w"] = T.glorot_uniform((1, F * hp, h), g, dtype).reshape(F * hp, h)
            p[f"cin.{k}.b"] = torch.zeros(h, dtype=dtype)
            hp = h
        # |.|: the pooled CIN maps are >= 0, a negative kernel would leave the ReLU of
        # xdeepfm.py:182 dead for every sample and the parity tests would not exercise CIN
        p["cin.out.w"] = T.glorot_uniform((sum(cin_layers), 1), g, dtype).abs()
        p["cin.out.b"] = torch.full((1,), 0.1, dtype=dtype)
    if model == "dcn":
        for l in range(cross_layers):
            p[f"cross.{l}.w"] = T.glorot_normal((F * D,), g, dtype)
            p[f"cross.{l}.b"] = T.glorot_normal((F * D,), g, dtype)
        _tower(p, "dnn", [F * D] + list(deep_layers), g, dtype, bn=True)
        p["head.w"] = T.glorot_uniform((deep_layers[-1] + F * D, 1), g, dtype)
        p["head.b"] = torch.zeros(1, dtype=dtype)
    if model == "din":
        p["i_item"] = torch.zeros(din_items, dtype=dtype)
        p["i_id"] = T.glorot_normal((din_items, D), g, dtype)
        p["i_cate"] = T.glorot_normal((din_cates, D), g, dtype)
        for name in ("att_iid", "att_cat"):
            _tower(p, name, [4 * D] + DIN_ATT_LAYERS + [1], g, dtype, bn=False)
        _tower(p, "mlp", [3 * D] + DIN_MLP_LAYERS, g, dtype, bn=False)
        p["mlp.out.w"] = T.glorot_uniform((DIN_MLP_LAYERS[-1], 1), g, dtype)
        p["mlp.out.b"] = torch.zeros(1, dtype=dtype)
    return p


# ------------------------------------------------------------------- sub-graphs
def gather(table, rows, sparse_grad=False):
    """input_layer over embedding columns = per-field gather, concat (K1/A3)."""
    if sparse_grad:
        return torch.nn.functional.embedding(rows, table, sparse=True)
    return table[rows]


def first_order(p, rows, mask=None):
    """dense(one-hot linear_net, 1, relu) == ReLU(sum_f w1[row_f] + b1)
    (fm/fm.py:117,120-121; deepfm/deepfm.py:87-91)."""
    w = p["w1"][rows]
    if mask is not None:
        w = w * mask
    return torch.relu(w.sum(1, keepdim=True) + p["b1"])


def fm_second_order(E):
    """0.5 * sum_d[(sum_f E)^2 - sum_f E^2], keep_dims (fm/fm.py:123-129)."""
    s = E.sum(1)
    return 0.5 * (s * s - (E * E).sum(1)).sum(1, keepdim=True)


def dnn_tower(p, x, n_layers, training, rate, masks, prefix="dnn", bn=True):
    """dense(relu) -> batch_normalization -> dropout, per layer
    (deepfm/deepfm.py:103-107, xdeepfm/xdeepfm.py:188-191, dcn/dcn.py:146-149)."""
    for l in range(n_layers):
        x = T.dense(x, p[f"{prefix}.{l}.w"], p[f"{prefix}.{l}.b"], relu=True)
        if bn:
            x = T.batch_norm(x, p[f"{prefix}.{l}.bn.gamma"], p[f"{prefix}.{l}.bn.beta"],
                             p[f"{prefix}.{l}.bn.mean"], p[f"{prefix}.{l}.bn.var"], training)
        x = T.dropout(x, rate, training, None if masks is None else masks[l])
    return x


def _n_layers(p, prefix):
    n = 0
    while f"{prefix}.{n}.w" in p:
        n += 1
    return n


def cin(p, X0):
    """xdeepfm/xdeepfm.py:135-182, literal op order: split over D, batched
    outer product, reshape to q = i*Hp + j, transpose, conv1d(k=1), bias, ReLU,
    transpose; every map goes to both the output and the next layer; sum over D."""
    B, m, D = X0.shape
    hidden = [X0]
    finals = []
    split0 = X0.permute(2, 0, 1).unsqueeze(-1)                    # [D,B,m,1]      :143
    k = 0
    while f"cin.{k}.w" in p:
        Xk = hidden[-1]
        splitk = Xk.permute(2, 0, 1).unsqueeze(-1)                # [D,B,Hp,1]     :146
        dot_m = split0 @ splitk.transpose(-1, -2)                 # [D,B,m,Hp]     :147
        dot_o = dot_m.reshape(D, B, m * Xk.shape[1])              #                :148-149
        dot = dot_o.permute(1, 0, 2)                              # [B,D,m*Hp]     :150
        out = dot @ p[f"cin.{k}.w"] + p[f"cin.{k}.b"]             # conv1d k=1     :156-163
        out = torch.relu(out).permute(0, 2, 1)                    # [B,H,D]        :166-167
        finals.append(out)
        hidden.append(out)
        k += 1
    result = torch.cat(finals, dim=1).sum(-1)                     # [B,sumH]       :180-181
    return T.dense(result, p["cin.out.w"], p["cin.out.b"], relu=True)   # :182


def dcn_cross(p, x0):
    """dcn/dcn.py:132-142: xl <- (xl . w) * x0 + xl + b."""
    xl = x0
    l = 0
    while f"cross.{l}.w" in p:
        xw = (xl * p[f"cross.{l}.w"]).sum(1, keepdim=True)
        xl = xw * x0 + xl + p[f"cross.{l}.b"]
        l += 1
    return xl


def din_attention(p, prefix, table, hist, query, training, rate, masks):
    """din/din.py:103-125 ``_attention``: no softmax, mask = id > 0."""
    B, P = hist.shape
    E = table.shape[1]
    dense_emb = table[hist]                                       # [B,P,E]
    mask = (hist > 0).to(table.dtype).unsqueeze(-1)
    h = dense_emb.reshape(-1, E)
    q = query.repeat(1, P).reshape(-1, E)
    a = torch.cat([h, q, h * q, h - q], dim=1)
    for l in range(len(DIN_ATT_LAYERS)):
        a = T.dense(a, p[f"{prefix}.{l}.w"], p[f"{prefix}.{l}.b"], relu=True)
        a = T.dropout(a, rate, training, None if masks is None else masks[l])
    L = len(DIN_ATT_LAYERS)
    w = T.dense(a, p[f"{prefix}.{L}.w"], p[f"{prefix}.{L}.b"]).reshape(B, P, 1)
    return (dense_emb * w * mask).sum(1)


# ------------------------------------------------------------------ model graphs
def _finish(logits, labels):
    out = {"logits": logits, "prob": torch.sigmoid(logits)}
    if labels is not None:
        z = labels.to(logits.dtype).reshape(logits.shape)
        out["loss"] = T.sigmoid_cross_entropy_with_logits(logits, z).mean()
    return out


def fm(p, rows, labels=None, training=False, sparse_grad=False, **_):
    """fm/fm.py:115-170."""
    E = gather(p["emb"], rows, sparse_grad)
    y1 = first_order(p, rows)
    y2 = fm_second_order(E)
    logits = T.dense(torch.cat([y1, y2], -1), p["head.w"], p["head.b"])       # [B,1]
    return _finish(logits, labels)


def deepfm(p, rows, labels=None, training=False, dropout=0.0, masks=None, sparse_grad=False, **_):
    """deepfm/deepfm.py:73-150 fed with the Criteo columns of fm/fm.py:47-97."""
    E = gather(p["emb"], rows, sparse_grad)
    y1 = first_order(p, rows)
    y2 = fm_second_order(E)
    h = dnn_tower(p, E.reshape(E.shape[0], -1), _n_layers(p, "dnn"), training, dropout, masks)
    y3 = T.dense(h, p["dnn.out.w"], p["dnn.out.b"], relu=True)
    logits = T.dense(torch.cat([y1, y2, y3], -1), p["head.w"], p["head.b"]).reshape(-1)  # [B]
    return _finish(logits, labels)


def xdeepfm(p, rows, logx, cat_mask, labels=None, training=False, dropout=0.0, masks=None,
            sparse_grad=False, **_):
    """xdeepfm/xdeepfm.py:123-233.  ``logx[B,13]``: the 13 numeric linear inputs
    (:82); ``cat_mask[F]``: 1 for the 26 indicator fields (:91).  The DNN branch
    reads ``emb_dnn`` when present - the second input_layer call (:185) creates
    a second set of tables [TF-sem]."""
    E = gather(p["emb"], rows, sparse_grad)
    lin = (p["w1"][rows] * cat_mask).sum(1, keepdim=True) + logx @ p["wnum"].reshape(-1, 1)
    linear_y = torch.relu(lin + p["b1"])                                        # :131
    cin_y = cin(p, E)
    Ed = gather(p["emb_dnn"], rows, sparse_grad) if "emb_dnn" in p else E       # :185
    h = dnn_tower(p, Ed.reshape(Ed.shape[0], -1), _n_layers(p, "dnn"), training, dropout, masks)
    dnn_y = T.dense(h, p["dnn.out.w"], p["dnn.out.b"], relu=True)
    logits = T.dense(torch.cat([linear_y, cin_y, dnn_y], -1), p["head.w"], p["head.b"])  # [B,1]
    return _finish(logits, labels)


def dcn(p, rows, labels=None, training=False, dropout=0.0, masks=None, sparse_grad=False, **_):
    """dcn/dcn.py:117-190 (linear_net is built but unused, :122,129-130)."""
    E = gather(p["emb"], rows, sparse_grad)
    x0 = E.reshape(E.shape[0], -1)
    xl = dcn_cross(p, x0)
    h = dnn_tower(p, x0, _n_layers(p, "dnn"), training, dropout, masks)
    logits = T.dense(torch.cat([h, xl], -1), p["head.w"], p["head.b"])          # [B,1]
    return _finish(logits, labels)


def din(p, i_id, i_cate, u_iid_seq, u_icat_seq, labels=None, training=False, dropout=0.0,
        masks=None, **_):
    """din/din.py:83-180.  ``masks`` = {"att_iid": [m0,m1], "att_cat": [...], "mlp": [m0,m1,m2]}."""
    masks = masks or {}
    i_b = p["i_item"][i_id]                                                     # :91
    pkg_emb = p["i_id"][i_id]
    pkgc_emb = p["i_cate"][i_cate]
    pkg_h = din_attention(p, "att_iid", p["i_id"], u_iid_seq, pkg_emb, training, dropout,
                          masks.get("att_iid"))
    pkgc_h = din_attention(p, "att_cat", p["i_cate"], u_icat_seq, pkgc_emb, training, dropout,
                           masks.get("att_cat"))
    net = torch.cat([pkg_emb, pkg_h, pkgc_h], dim=1)                            # :131
    net = dnn_tower(p, net, len(DIN_MLP_LAYERS), training, dropout, masks.get("mlp"),
                    prefix="mlp", bn=False)
    logits = T.dense(net, p["mlp.out.w"], p["mlp.out.b"]).reshape(-1) + i_b     # :139-140
    return _finish(logits, labels)


MODELS = {"fm": fm, "deepfm": deepfm, "xdeepfm": xdeepfm, "dcn": dcn, "din": din}


def loss_and_grads(model: str, p: dict, batch: dict, **kw):
    """Forward in train mode + autograd.  Returns (outputs, {name: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and
                                                   not k.endswith((".bn.mean", ".bn.var")))
              for k, v in p.items()}
    out = MODELS[model](leaves, **batch, training=True, **kw)
    out["loss"].backward()
    grads = {}
    for k, v in leaves.items():
        if v.grad is None:
            continue
        grads[k] = v.grad.to_dense() if v.grad.is_sparse else v.grad
    return {k: v.detach() for k, v in out.items()}, grads

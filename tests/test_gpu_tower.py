"""GPU parity, kernel level, for the dense-tower C ABI (ctr_tower_*): every GEMM flavour
(tcgen05 path and mma.sync path, selected with CTR_TOWER_TC) against a float64 torch
restatement of deepfm/deepfm.py:100-108 on the same inputs.  Tolerance: 3xTF32 keeps fp32-class
accuracy, so 2e-5 relative to the largest output of each tensor."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(tc):
    os.environ["CTR_TOWER_TC"] = "1" if tc else "0"


@pytest.fixture(autouse=True)
def _restore_env():
    old = os.environ.get("CTR_TOWER_TC")
    yield
    if old is None:
        os.environ.pop("CTR_TOWER_TC", None)
    else:
        os.environ["CTR_TOWER_TC"] = old


def _close(got, want, tol=2e-5):
    want = want.to(torch.float64).to(got.device)
    err = (got.to(torch.float64) - want).abs().max().item()
    ref = max(want.abs().max().item(), 1e-6)
    assert err <= tol * ref, "max err %.3e vs scale %.3e" % (err, ref)


SHAPES = [(4096, 624, 100), (300, 624, 100), (512, 128, 64), (1024, 256, 256), (4096, 100, 100),
          (2048, 416, 36)]


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("B,K,N", SHAPES)
def test_layer_fwd_plain(cuda, tc, B, K, N):
    from recsys_b200 import _lib, ops
    _env(tc)
    g = torch.Generator(device="cpu").manual_seed(B + K + N)
    X = torch.randn(B, K, generator=g).to(cuda)
    W = (torch.randn(K, N, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    out = torch.full((B, N), float("nan"), device=cuda)
    stats = torch.zeros(2, N, device=cuda)
    ops._call("ctr_tower_layer_fwd", X.data_ptr(), K, K, None, W.data_ptr(), b.data_ptr(), N,
              out.data_ptr(), N, stats.data_ptr(), 1, B, ops._stream())
    want = torch.relu(X.double() @ W.double() + b.double())
    _close(out, want)
    _close(stats[0], want.sum(0), 3e-5)
    _close(stats[1], (want * want).sum(0), 3e-5)
    # no ReLU, no stats
    out2 = torch.empty(B, N, device=cuda)
    ops._call("ctr_tower_layer_fwd", X.data_ptr(), K, K, None, W.data_ptr(), None, N,
              out2.data_ptr(), N, None, 0, B, ops._stream())
    _close(out2, X.double() @ W.double())


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("B,K,N", SHAPES)
def test_layer_bwd_plain(cuda, tc, B, K, N):
    """dpre from a BN gradient source, then dX = dpre . W^T and dW = X^T . dpre, db = colsum."""
    from recsys_b200 import _lib, ops
    _env(tc)
    g = torch.Generator(device="cpu").manual_seed(7 * B + K + N)
    X = torch.randn(B, K, generator=g).to(cuda)
    W = (torch.randn(K, N, generator=g) / K ** 0.5).to(cuda)
    a = torch.relu(torch.randn(B, N, generator=g)).to(cuda)            # stored post-ReLU output
    dn = torch.randn(B, N, generator=g).to(cuda)                       # gradient after the BN
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda)
    sums = torch.stack([a.sum(0), (a * a).sum(0)]).contiguous()
    eps = 1e-3
    a64, dn64 = a.double(), dn.double()
    mu = a64.mean(0)
    var = (a64 * a64).mean(0) - mu * mu
    rstd = 1 / torch.sqrt(var + eps)
    xhat = (a64 - mu) * rstd
    dbeta, dgamma = dn64.sum(0), (dn64 * xhat).sum(0)
    gbn = gamma.double() * rstd * (dn64 - dbeta / B - xhat * dgamma / B)
    dpre_want = gbn * (a64 > 0)

    gs = _lib.GradSrc()
    gs.G, gs.ldg, gs.a, gs.lda, gs.kind, gs.train, gs.eps = dn.data_ptr(), N, a.data_ptr(), N, 1, 1, eps
    dbeta32, dgamma32 = dbeta.float().contiguous(), dgamma.float().contiguous()
    gs.sums, gs.gamma = sums.data_ptr(), gamma.data_ptr()
    gs.dbeta, gs.dgamma = dbeta32.data_ptr(), dgamma32.data_ptr()
    dpre = torch.empty(B, N, device=cuda)
    db = torch.zeros(N, device=cuda)
    ops._call("ctr_tower_dpre", C.byref(gs), N, dpre.data_ptr(), N, db.data_ptr(), B, ops._stream())
    _close(dpre, dpre_want, 1e-4)
    _close(db, dpre_want.sum(0), 1e-4)

    g2 = _lib.GradSrc()
    g2.G, g2.ldg, g2.kind, g2.train, g2.eps = dpre.data_ptr(), N, 2, 1, eps
    dX = torch.full((B, K), float("nan"), device=cuda)
    ops._call("ctr_tower_layer_bwd_data", C.byref(g2), N, W.data_ptr(), K, None, None, dX.data_ptr(),
              K, None, None, B, ops._stream())
    _close(dX, dpre.double() @ W.double().t())
    dW = torch.zeros(K, N, device=cuda)
    db2 = torch.zeros(N, device=cuda)
    ops._call("ctr_tower_layer_bwd_weights", X.data_ptr(), K, K, None, C.byref(g2), N, dW.data_ptr(),
              db2.data_ptr(), B, ops._stream())
    _close(dW, X.double().t() @ dpre.double())
    _close(db2, dpre.double().sum(0), 1e-5)
    # kind 1 straight into the GEMMs (dpre re-derived inside the loaders) must agree too
    dX1 = torch.empty(B, K, device=cuda)
    ops._call("ctr_tower_layer_bwd_data", C.byref(gs), N, W.data_ptr(), K, None, None, dX1.data_ptr(),
              K, None, None, B, ops._stream())
    _close(dX1, dpre_want @ W.double().t(), 1e-4)
    torch.cuda.synchronize()


# ------------------------------------------------------------------ ctr_tower_mid (one launch)
def _masks(tw, B, dev):
    """Keep masks of the current step for every hidden layer (identity-BN trick, see
    test_gpu_models._dropout_masks)."""
    from recsys_b200 import _lib
    lib = _lib.load()
    out = []
    for l, H in enumerate(tw.sizes[1:]):
        ones = torch.ones(B, H, device=dev)
        res = torch.empty_like(ones)
        zero, one = torch.zeros(H, device=dev), torch.ones(H, device=dev)
        var = one - 1e-3
        d = _lib.BnDrop()
        d.sums, d.mean, d.var = None, zero.data_ptr(), var.data_ptr()
        d.gamma, d.beta, d.state = one.data_ptr(), zero.data_ptr(), tw.adam.state_ptr
        d.eps, d.p_drop, d.seed, d.layer, d.enabled = 1e-3, tw.dropout, tw.seed, l, 1
        rc = lib.ctr_bn_drop_apply(ones.data_ptr(), H, C.byref(d), res.data_ptr(), B,
                                   torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        out.append((res > 0).double())
    return out


def _tower_head_ref(P, X, zs, labels, sizes, masks, p, training, gates=None):
    """float64 restatement of deepfm/deepfm.py:100-129 (dense-relu-BN-dropout stack, dense(1,relu),
    head over [relu(z0 + b1), z1.., y], mean sigmoid cross entropy).

    ``gates``: the hidden ReLUs' on/off pattern taken from the kernel's own activations.  A ReLU
    whose pre-activation is within fp32 rounding of 0 may gate differently in fp32 and fp64 (and,
    because the BN column sums are accumulated with atomics, from run to run); that changes the
    forward by ~1e-7 but that row's gradient by O(1), which is a property of ReLU, not a kernel
    error.  With the kernel's gates the comparison is between smooth functions."""
    h = X
    for l in range(len(sizes) - 1):
        pre = h @ P["t.%d.w" % l] + P["t.%d.b" % l]
        a = torch.relu(pre) if gates is None else pre * gates[l]      # gates[-1]: the output ReLU
        if training:
            mu, var = a.mean(0), a.var(0, unbiased=False)
        else:
            mu, var = P["t.%d.bn.mean" % l], P["t.%d.bn.var" % l]
        h = (a - mu) / torch.sqrt(var + 1e-3) * P["t.%d.bn.gamma" % l] + P["t.%d.bn.beta" % l]
        if training and p > 0:
            h = h * masks[l] / (1 - p)
    pre_y = (h @ P["t.out.w"] + P["t.out.b"]).reshape(-1)
    y = torch.relu(pre_y) if gates is None else pre_y * gates[-1]
    cols = [torch.relu(zs[0] + P["b1"])] + list(zs[1:]) + [y]
    logit = torch.stack(cols, 1) @ P["head.w"].reshape(-1) + P["head.b"]
    loss = (torch.clamp(logit, min=0) - logit * labels + torch.log1p(torch.exp(-logit.abs()))).mean()
    return loss, logit


@pytest.mark.parametrize("presplit", [False, True])
@pytest.mark.parametrize("B,sizes,nz,p", [(4096, [624, 100, 100], 2, 0.5), (1000, [624, 100, 100], 2, 0.0),
                                           (333, [96, 32, 16], 2, 0.5), (517, [64, 128], 1, 0.0),
                                           (6000, [128, 64, 32, 16, 8], 2, 0.3)])
def test_tower_mid_matches_float64(cuda, B, sizes, nz, p, presplit):
    from recsys_b200 import ops
    torch.manual_seed(B)
    shapes = {"b1": (1,), "head.w": (nz + 1, 1), "head.b": (1,), "t.out.w": (sizes[-1], 1), "t.out.b": (1,)}
    for l, (i, o) in enumerate(zip(sizes[:-1], sizes[1:])):
        shapes.update({"t.%d.w" % l: (i, o), "t.%d.b" % l: (o,), "t.%d.bn.gamma" % l: (o,),
                       "t.%d.bn.beta" % l: (o,), "t.%d.bn.mean" % l: (o,), "t.%d.bn.var" % l: (o,)})
    frozen = [n for n in shapes if n.endswith((".bn.mean", ".bn.var"))]
    dense = ops.DenseParams(shapes, cuda, frozen=frozen)
    with torch.no_grad():
        for n in dense.names:
            v = dense[n]
            if n.endswith(".bn.var"):
                v.copy_(torch.rand_like(v) + 0.5)
            elif n.endswith(".w"):
                v.copy_(torch.randn_like(v) * (2.0 / v.shape[0]) ** 0.5)
            else:
                v.copy_(torch.randn_like(v) * 0.3 + (1.0 if n.endswith("gamma") else 0.0))
    adam = ops.TFAdamState(device=cuda)
    adam.next_lr_t()                                   # step counter 1 selects the dropout stream
    tw = ops.FusedTower(dense, "t", sizes, True, p, adam, seed=11)
    assert tw.use_mid
    X = torch.randn(B, sizes[0], device=cuda, requires_grad=True)
    zs = [torch.randn(B, device=cuda, requires_grad=True) for _ in range(nz)]
    labels = (torch.rand(B, device=cuda) < 0.3).float()
    masks = _masks(tw, B, cuda) if p > 0 else None
    P64 = {n: dense[n].detach().double().requires_grad_(n not in frozen) for n in dense.names}
    X64 = X.detach().double().requires_grad_(True)
    zs64 = [z.detach().double().requires_grad_(True) for z in zs]
    for training in (False, True):
        X_lo = ops.split_lo(X.detach()) if presplit else None   # first-layer GEMMs on pre-split operands
        loss, logits, prob = ops.tower_head(tw, X, zs, labels, training=training, X_lo=X_lo)
        torch.cuda.synchronize()
        gates = [(a > 0).double() for a in tw.last_acts] + [(tw.last_y > 0).double()] \
            if training else None
        loss64, logit64 = _tower_head_ref(P64, X64, zs64, labels.double(), sizes, masks, p, training,
                                          gates)
        _close(logits, logit64.detach(), 1e-4)
        _close(prob, torch.sigmoid(logit64.detach()), 1e-4)
        assert abs(float(loss) - float(loss64)) <= 1e-5 * max(1.0, abs(float(loss64)))
    loss64.backward()
    loss.backward()
    tw.join()
    torch.cuda.synchronize()
    _close(X.grad, X64.grad, 1e-4)
    for z, z64 in zip(zs, zs64):
        _close(z.grad, z64.grad, 1e-4)
    for n in dense.names:
        if n not in frozen:
            _close(dense[n].grad, P64[n].grad.reshape(dense[n].shape), 2e-4)


def test_tower_mid_agrees_with_per_layer_kernels(cuda):
    """Same seeds, same dropout stream: the one-launch path and the per-layer path (CTR_TOWER_MID=0)
    of DeepFM produce the same loss and the same gradients."""
    import numpy as np
    import make_golden as mg
    from oracle import models as om
    from test_gpu_models import _build, _features_to_torch
    from recsys_b200.deepfm import deepfm
    spec = mg.small_spec()
    p64 = om.init_params("deepfm", spec.total_rows, deep_layers=(100, 100), seed=5)
    feats, batch = mg.model_batch("deepfm", 2048, 4, spec)
    res = {}
    for mid in ("1", "0"):
        os.environ["CTR_TOWER_MID"] = mid
        try:
            m, params = _build("deepfm", spec, cuda, deep_layers="100,100", dropout=0.5)
            m.load_state(p64)
            sp = deepfm.model_fn(_features_to_torch(feats), batch["labels"], "train", params)
            m.backward(m.last["loss"])
            torch.cuda.synchronize()
            g = {k: v.detach().clone() for k, v in m.dense_grads().items()}
            g["emb"] = m.emb.dtable.detach().clone()
            res[mid] = (float(sp.loss), g)
        finally:
            os.environ.pop("CTR_TOWER_MID", None)
    assert abs(res["1"][0] - res["0"][0]) <= 1e-6
    # relative Frobenius error: one ReLU gating differently within fp32 rounding (the BN sums are
    # atomics) moves a single row's gradient, which a max-norm check would flag
    for k, v in res["0"][1].items():
        d = (res["1"][1][k].double() - v.double()).norm().item()
        assert d <= 2e-3 * max(v.double().norm().item(), 1e-12), (k, d)


@pytest.mark.parametrize("B,H,W", [(4096, 100, 624), (77, 32, 64), (1000, 16, 1280)])
def test_dcn_head_matches_float64(cuda, B, H, W):
    """ctr_dcn_head (dcn/dcn.py:151-153,166-169): logits, prob, mean BCE and every gradient in
    one launch, against float64 autograd."""
    from recsys_b200 import ops
    g = torch.Generator().manual_seed(B + H + W)
    h = torch.randn(B, H, generator=g)
    xl = torch.randn(B, W, generator=g)
    w = torch.randn(H + W, generator=g) / (H + W) ** 0.5
    hb = torch.tensor([0.3])
    z = (torch.rand(B, generator=g) < 0.3).float()
    d = {k: v.to(cuda).contiguous() for k, v in dict(h=h, xl=xl, w=w, hb=hb, z=z).items()}
    logits, prob = torch.empty(B, device=cuda), torch.empty(B, device=cuda)
    loss = torch.zeros((), device=cuda)
    dh, dxl = torch.empty(B, H, device=cuda), torch.empty(B, W, device=cuda)
    dw, dhb = torch.zeros(H + W, device=cuda), torch.zeros(1, device=cuda)
    ops._call("ctr_dcn_head", d["h"].data_ptr(), H, d["xl"].data_ptr(), W, d["w"].data_ptr(),
              d["hb"].data_ptr(), d["z"].data_ptr(), B, logits.data_ptr(), prob.data_ptr(),
              loss.data_ptr(), dh.data_ptr(), dxl.data_ptr(), dw.data_ptr(), dhb.data_ptr(),
              1.0 / B, ops._stream())
    h64, xl64 = h.double().requires_grad_(True), xl.double().requires_grad_(True)
    w64, hb64 = w.double().requires_grad_(True), hb.double().requires_grad_(True)
    lg = torch.cat([h64, xl64], 1) @ w64 + hb64
    ls = torch.nn.functional.binary_cross_entropy_with_logits(lg, z.double())
    ls.backward()
    _close(logits, lg.detach(), 1e-5)
    _close(prob, torch.sigmoid(lg.detach()), 1e-5)
    assert abs(loss.item() - ls.item()) <= 1e-6
    _close(dh, h64.grad, 1e-5)
    _close(dxl, xl64.grad, 1e-5)
    _close(dw, w64.grad, 2e-5)
    _close(dhb, hb64.grad, 2e-5)
    # inference flavour: no gradient outputs
    loss2 = torch.zeros((), device=cuda)
    ops._call("ctr_dcn_head", d["h"].data_ptr(), H, d["xl"].data_ptr(), W, d["w"].data_ptr(),
              d["hb"].data_ptr(), d["z"].data_ptr(), B, logits.data_ptr(), None, loss2.data_ptr(),
              None, None, None, None, 1.0 / B, ops._stream())
    assert abs(loss2.item() - ls.item()) <= 1e-6


@pytest.mark.parametrize("B,K,p", [(4096, 100, 0.5), (130, 36, 0.0), (512, 256, 0.3)])
def test_bn_drop_apply_and_its_backward(cuda, B, K, p):
    """ctr_bn_drop_apply / ctr_bn_drop_apply_bwd: out = dropout(BN_train(A)); the backward hands
    dn = dout * keep and the BN column sums to a kind-1 gradient source, whose dpre must equal
    float64 autograd through relu -> batch-norm (batch statistics) -> the same dropout mask."""
    from recsys_b200 import _lib, ops
    g = torch.Generator().manual_seed(B + K)
    pre = torch.randn(B, K, generator=g)
    A = torch.relu(pre).to(cuda)
    gamma = (torch.rand(K, generator=g) + 0.5).to(cuda)
    beta = torch.randn(K, generator=g).to(cuda)
    dout = torch.randn(B, K, generator=g).to(cuda)
    sums = torch.stack([A.sum(0), (A * A).sum(0)]).contiguous()
    d = _lib.BnDrop()
    d.sums, d.gamma, d.beta, d.state = sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None
    d.eps, d.p_drop, d.seed, d.layer, d.enabled = 1e-3, p, 5, 1, 1
    out = torch.empty(B, K, device=cuda)
    ops._call("ctr_bn_drop_apply", A.data_ptr(), K, C.byref(d), out.data_ptr(), B, ops._stream())
    a64 = pre.double().requires_grad_(True)
    r = torch.relu(a64)
    mu, var = r.mean(0), r.var(0, unbiased=False)
    bn = (r - mu) / torch.sqrt(var + 1e-3) * gamma.cpu().double() + beta.cpu().double()
    # the mask the kernel used: where the output is exactly zero although bn is not
    keep = torch.ones(B, K, dtype=torch.float64)
    if p > 0:
        keep = (out.cpu().double() != 0).double()
        assert abs(keep.mean().item() - (1 - p)) < 0.03
    want = bn * keep / (1 - p)
    _close(out, want.detach(), 2e-5)
    want.backward(dout.cpu().double())
    dn = torch.empty(B, K, device=cuda)
    dbeta, dgamma = torch.zeros(K, device=cuda), torch.zeros(K, device=cuda)
    ops._call("ctr_bn_drop_apply_bwd", dout.data_ptr(), K, A.data_ptr(), K, C.byref(d),
              dn.data_ptr(), dbeta.data_ptr(), dgamma.data_ptr(), B, ops._stream())
    _close(dn, dout.cpu().double() * keep / (1 - p), 1e-6)
    gs = _lib.GradSrc()
    gs.G, gs.ldg, gs.a, gs.lda, gs.kind, gs.train, gs.eps = dn.data_ptr(), K, A.data_ptr(), K, 1, 1, 1e-3
    gs.sums, gs.gamma = sums.data_ptr(), gamma.data_ptr()
    gs.dbeta, gs.dgamma = dbeta.data_ptr(), dgamma.data_ptr()
    dpre = torch.empty(B, K, device=cuda)
    ops._call("ctr_tower_dpre", C.byref(gs), K, dpre.data_ptr(), K, None, B, ops._stream())
    _close(dpre, a64.grad, 5e-5)

"""Debug harness for the tcgen05 tower GEMMs: structured inputs, prints error summaries."""
import ctypes as C
import os
import sys
import torch
from recsys_b200 import _lib, ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def summary(name, got, want):
    want = want.double()
    got = got.double()
    err = (got - want).abs()
    print("%-28s max|err| %.3e  scale %.3e  nan %d  got[0,:6] %s want[0,:6] %s" % (
        name, err.max().item(), want.abs().max().item(), int(torch.isnan(got).sum()),
        [round(v, 3) for v in got[0, :6].tolist()], [round(v, 3) for v in want[0, :6].tolist()]))
    sys.stdout.flush()


def fwd(X, W, tag):
    B, K = X.shape
    N = W.shape[1]
    out = torch.full((B, N), float("nan"), device=dev)
    ops._call("ctr_tower_layer_fwd", X.data_ptr(), K, K, None, W.data_ptr(), None, N,
              out.data_ptr(), N, None, 0, B, ops._stream())
    torch.cuda.synchronize()
    summary("fwd " + tag, out, X.double() @ W.double())
    return out


def bwd_data(D, W, tag):
    B, N = D.shape
    K = W.shape[0]
    g2 = _lib.GradSrc()
    g2.G, g2.ldg, g2.kind, g2.train, g2.eps = D.data_ptr(), N, 2, 1, 1e-3
    dX = torch.full((B, K), float("nan"), device=dev)
    ops._call("ctr_tower_layer_bwd_data", C.byref(g2), N, W.data_ptr(), K, None, None, dX.data_ptr(),
              K, None, None, B, ops._stream())
    torch.cuda.synchronize()
    summary("bwd_data " + tag, dX, D.double() @ W.double().t())


def bwd_w(X, D, tag):
    B, K = X.shape
    N = D.shape[1]
    g2 = _lib.GradSrc()
    g2.G, g2.ldg, g2.kind, g2.train, g2.eps = D.data_ptr(), N, 2, 1, 1e-3
    dW = torch.zeros(K, N, device=dev)
    ops._call("ctr_tower_layer_bwd_weights", X.data_ptr(), K, K, None, C.byref(g2), N, dW.data_ptr(),
              None, B, ops._stream())
    torch.cuda.synchronize()
    summary("bwd_w " + tag, dW, X.double().t() @ D.double())


B, K, N = 4096, 624, 100
for passes in ("1", "3"):
    os.environ["CTR_TCG_PASSES"] = passes
    print("== passes", passes)
    X = torch.randn(B, K, device=dev)
    W = torch.randn(K, N, device=dev) / K ** 0.5
    D = torch.randn(B, N, device=dev)
    bwd_data(D, W, "rand")
    fwd(X, W, "rand")
    bwd_w(X, D, "rand")
    ones = torch.ones(B, K, device=dev)
    Wn = (torch.arange(N, device=dev, dtype=torch.float32) + 1).repeat(K, 1).contiguous()
    fwd(ones, Wn, "X=1 W=n+1")
    Wk = torch.zeros(K, N, device=dev); Wk[0, :] = 1.0
    fwd(X, Wk, "W=e_k0")
    Wk = torch.zeros(K, N, device=dev); Wk[5, :] = 1.0
    fwd(X, Wk, "W=e_k5")
    Wk = torch.zeros(K, N, device=dev); Wk[40, 3] = 1.0
    o = fwd(X, Wk, "W=e_k40,n3")
    nz = (o[0].abs() > 1e-6).nonzero().flatten().tolist()
    print("   nonzero cols row0:", nz[:10], "x[0,40]=%.4f" % X[0, 40].item(), "vals", [round(o[0, c].item(), 4) for c in nz[:10]])

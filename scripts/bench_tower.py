"""Micro-timings of the tower ABI calls (CUDA events, L2-warm and L2-flushed)."""
import ctypes as C
import os
import sys
import torch
from recsys_b200 import _lib, ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, K, N = int(os.environ.get("TB", 4096)), int(os.environ.get("TK", 624)), int(os.environ.get("TN", 100))
X = torch.randn(B, K, device=dev)
W = torch.randn(K, N, device=dev) / K ** 0.5
bias = torch.randn(N, device=dev)
D = torch.randn(B, N, device=dev)
a = torch.relu(torch.randn(B, N, device=dev))
out = torch.empty(B, N, device=dev)
stats = torch.zeros(2, N, device=dev)
dX = torch.empty(B, K, device=dev)
dW = torch.zeros(K, N, device=dev)
db = torch.zeros(N, device=dev)
dpre = torch.empty(B, N, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g2 = _lib.GradSrc()
g2.G, g2.ldg, g2.kind, g2.train, g2.eps = D.data_ptr(), N, 2, 1, 1e-3
g0 = _lib.GradSrc()
g0.G, g0.ldg, g0.a, g0.lda, g0.kind, g0.train, g0.eps = D.data_ptr(), N, a.data_ptr(), N, 0, 1, 1e-3
s = ops._stream


def t(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / reps * 1e3
    cold = []
    for _ in range(5):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1) * 1e3)
    print("%-34s warm %7.1f us   cold %7.1f us" % (name, warm, sorted(cold)[2]))
    sys.stdout.flush()


def fwd():
    ops._call("ctr_tower_layer_fwd", X.data_ptr(), K, K, None, W.data_ptr(), bias.data_ptr(), N,
              out.data_ptr(), N, stats.data_ptr(), 1, B, s())


def bwd_d():
    ops._call("ctr_tower_layer_bwd_data", C.byref(g2), N, W.data_ptr(), K, None, None, dX.data_ptr(),
              K, None, None, B, s())


def bwd_w():
    ops._call("ctr_tower_layer_bwd_weights", X.data_ptr(), K, K, None, C.byref(g2), N, dW.data_ptr(),
              None, B, s())


def dp():
    ops._call("ctr_tower_dpre", C.byref(g0), N, dpre.data_ptr(), N, db.data_ptr(), B, s())


print("B=%d K=%d N=%d" % (B, K, N))
for tc in ("0", "1"):
    os.environ["CTR_TOWER_TC"] = tc
    t("fwd tc=%s" % tc, fwd)
    t("bwd_data tc=%s" % tc, bwd_d)
    t("bwd_weights tc=%s" % tc, bwd_w)
t("dpre", dp)
for nt in (1, 2, 4):
    os.environ["CTR_TCG_NTILES"] = str(nt)
    t("fwd ntiles=%d" % nt, fwd)
for nt in (3, 4, 5, 8):
    os.environ["CTR_TCG_NTILES"] = str(nt)
    t("bwd_data ntiles=%d" % nt, bwd_d)
os.environ.pop("CTR_TCG_NTILES")
for sp in (8, 16, 29):
    os.environ["CTR_TCG_SPLITS"] = str(sp)
    t("bwd_weights splits=%d" % sp, bwd_w)
os.environ.pop("CTR_TCG_SPLITS")
for stg in (2, 3, 4, 6):
    os.environ["CTR_TCG_STAGES"] = str(stg)
    t("fwd stages=%d" % stg, fwd)
os.environ.pop("CTR_TCG_STAGES")
os.environ["CTR_TCG_PASSES"] = "1"
t("fwd 1-pass", fwd)
t("bwd_data 1-pass", bwd_d)
t("bwd_weights 1-pass", bwd_w)

"""Micro-benchmarks (GPU): (1) ctr_tower_mid phase profile at the DeepFM bench shape,
(2) embed fwd / bwd / adam_rows time against the table size (TLB / DRAM-page reach)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from recsys_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def ev_time(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


def mid_profile(B=4096, sizes=(624, 100, 100), p=0.5):
    sizes = list(sizes)
    shapes = {"b1": (1,), "head.w": (3, 1), "head.b": (1,), "t.out.w": (sizes[-1], 1), "t.out.b": (1,)}
    for l, (i, o) in enumerate(zip(sizes[:-1], sizes[1:])):
        shapes.update({"t.%d.w" % l: (i, o), "t.%d.b" % l: (o,), "t.%d.bn.gamma" % l: (o,),
                       "t.%d.bn.beta" % l: (o,), "t.%d.bn.mean" % l: (o,), "t.%d.bn.var" % l: (o,)})
    frozen = [n for n in shapes if n.endswith((".bn.mean", ".bn.var"))]
    dense = ops.DenseParams(shapes, dev, frozen=frozen)
    with torch.no_grad():
        for n in dense.names:
            v = dense[n]
            if n.endswith(".w"):
                v.copy_(torch.randn_like(v) * (2.0 / v.shape[0]) ** 0.5)
            elif n.endswith(("gamma", "var")):
                v.fill_(1.0)
    adam = ops.TFAdamState(device=dev)
    adam.next_lr_t()
    tw = ops.FusedTower(dense, "t", sizes, True, p, adam, seed=1)
    tw.timing = torch.zeros(8, dtype=torch.int64, device=dev)
    X = torch.randn(B, sizes[0], device=dev, requires_grad=True)
    zs = [torch.randn(B, device=dev, requires_grad=True) for _ in range(2)]
    labels = (torch.rand(B, device=dev) < 0.3).float()

    def fwd_bwd():
        loss, _, _ = ops.tower_head(tw, X, zs, labels, training=True)
        loss.backward()
        tw.join()

    def fwd_only():
        with torch.no_grad():
            ops.tower_head(tw, X, zs, labels, training=True)

    t_all = ev_time(fwd_bwd)
    t_fwd = ev_time(fwd_only)
    stamps = []
    for _ in range(10):
        fwd_only()
        torch.cuda.synchronize()
        t = tw.timing.cpu().tolist()
        stamps.append([(t[i + 1] - t[i]) / 1e3 for i in range(7)])
    med = [sorted(s[i] for s in stamps)[len(stamps) // 2] for i in range(7)]
    print("tower_head B=%d sizes=%s p=%g: eager fwd+bwd %.1f us, layer0 GEMM + mid (no backward GEMMs) %.1f us"
          % (B, sizes, p, t_all, t_fwd))
    print("  mid phases us: F(hidden fwd) %.1f | barrier %.1f | O(out+head) %.1f | barrier %.1f | "
          "B(hidden bwd) %.1f | barrier %.1f | D0 %.1f | total %.1f" % (*med, sum(med)))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fwd_only()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fwd_only()
    t_g = ev_time(lambda: g.replay())
    print("  graph replay of {zero ws, layer0 GEMM, mid}: %.1f us" % t_g)


def table_sweep(B=4096, F=39, D=16):
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for R in (1 << 20, 1 << 22, 1 << 24, 33762673, 1 << 26):
        table = torch.randn(R, D, device=dev)
        w1 = torch.randn(R, device=dev)
        dt, m, v = torch.zeros_like(table), torch.zeros_like(table), torch.zeros_like(table)
        dw1, m1, v1 = torch.zeros_like(w1), torch.zeros_like(w1), torch.zeros_like(w1)
        claim = torch.zeros(R, dtype=torch.int32, device=dev)
        nb = 16
        rows = [torch.randint(0, R, (B, F), device=dev, dtype=torch.int32) for _ in range(nb)]
        E = torch.empty(B, F * D, device=dev)
        S = torch.empty(B, D, device=dev)
        y1 = torch.empty(B, device=dev)
        y2 = torch.empty(B, device=dev)
        dE = torch.randn(B, F * D, device=dev)
        dy = torch.randn(B, device=dev)
        offs = (torch.arange(F + 1, dtype=torch.int64) * (R // F)).tolist()
        offs[-1] = R
        offs_c = (C.c_int64 * (F + 1))(*offs)
        mask = (1 << F) - 1
        it = [0]
        tag = [0]

        def fwd():
            r = rows[it[0] % nb]
            it[0] += 1
            lib.ctr_embed_fwd(table.data_ptr(), w1.data_ptr(), r.data_ptr(), B, F, D, mask, E.data_ptr(),
                              S.data_ptr(), y1.data_ptr(), y2.data_ptr(), None, None, 0, None, None, 0, 0, st)

        def bwd():
            r = rows[it[0] % nb]
            it[0] += 1
            lib.ctr_embed_bwd(r.data_ptr(), dE.data_ptr(), E.data_ptr(), table.data_ptr(), S.data_ptr(),
                              dy.data_ptr(), dy.data_ptr(), mask, offs_c, B, F, D, dt.data_ptr(),
                              dw1.data_ptr(), 0, 0, st)

        def adam():
            r = rows[it[0] % nb]
            it[0] += 1
            tag[0] += 1
            lib.ctr_adam_rows(r.data_ptr(), B * F, D, table.data_ptr(), m.data_ptr(), v.data_ptr(),
                              dt.data_ptr(), w1.data_ptr(), m1.data_ptr(), v1.data_ptr(), dw1.data_ptr(),
                              claim.data_ptr(), tag[0], 1e-3, 0.9, 0.999, 1e-8, None, 0, 0, 0, st)

        print("R=%9d rows (%.2f GB table): fwd %.1f us  bwd %.1f us  adam_rows %.1f us" % (
            R, R * D * 4 / 1e9, ev_time(fwd, 64), ev_time(bwd, 64), ev_time(adam, 64)))
        del table, w1, dt, m, v, dw1, m1, v1, claim
        torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["mid", "sweep"]
    if "mid" in which:
        mid_profile()
        mid_profile(p=0.0)
    if "sweep" in which:
        table_sweep()

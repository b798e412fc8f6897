"""Micro-benchmark (GPU): ctr_embed_tower_fwd alone on the R-full Criteo table - per-launch time
(8 distinct batches in a CUDA graph) and the phase stamps of CTA (0,0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from recsys_b200 import _lib, ops  # noqa: E402
from recsys_b200 import feature_column as fc  # noqa: E402
from recsys_b200.data import SyntheticCriteo  # noqa: E402
from recsys_b200.fm import fm  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()


def main():
    B, N = int(os.environ.get("B", 4096)), 100
    _, cols = fm.build_feature_columns(16, full_cardinality=os.environ.get("TABLE", "full") == "full")
    lay = fc.layout(cols)
    F, D = lay.F, 16
    emb = ops.FieldEmbedding(lay, dev, with_w1=True, w1_fields=(1 << F) - 1, adam_mode="lazy")
    pipe = ops.IdPipeline(lay, dev)
    host = SyntheticCriteo(lay, B, 8, dist="uniform", seed=0, device=None).batches
    packs = [pipe.pack(ops.PackedFeatures(f.cont.to(dev), f.cat.to(dev), f.cont_keys, f.cat_keys))
             for f, _ in host]
    W0 = torch.randn(F * D, N, device=dev) * 0.05
    W0_lo = ops.split_lo(W0)
    b0 = torch.zeros(N, device=dev)
    p = ops._p
    rows = torch.empty(B, F, dtype=torch.int32, device=dev)
    E = torch.empty(B, F * D, device=dev)
    E_lo = torch.empty_like(E)
    S = torch.empty(B, D, device=dev)
    y1, y2 = torch.empty(B, device=dev), torch.empty(B, device=dev)
    act0 = torch.empty(B, N, device=dev)
    parts = torch.empty((B + 127) // 128, 2, N, device=dev)
    zbuf = torch.empty(404, device=dev)
    s = torch.cuda.Stream()

    def fused(i):
        cont, cat = packs[i % 8]
        rc = lib.ctr_embed_tower_fwd(p(emb.table), p(emb.w1), p(cont), len(pipe.cont_keys), p(cat),
                                     len(pipe.cat_keys), p(pipe.fields_dev), p(pipe.bnd_dev), pipe.n_bnd, None, p(rows),
                                     p(pipe.status), B, F, D, emb.w1_fields, p(E), p(E_lo), p(S), p(y1),
                                     p(y2), emb.ld, emb.ld1, p(W0), p(W0_lo), p(b0), N, p(act0), p(parts),
                                     p(zbuf), zbuf.numel(), s.cuda_stream)
        assert rc == 0, _lib.last_error()

    def unfused(i):
        cont, cat = packs[i % 8]
        lib.ctr_embed_fwd_raw(p(emb.table), p(emb.w1), p(cont), len(pipe.cont_keys), p(cat),
                              len(pipe.cat_keys), p(pipe.fields_dev), p(pipe.bnd_dev), pipe.n_bnd,
                              p(rows), None, p(pipe.status), B, F, D, emb.w1_fields, p(E), p(S), p(y1),
                              p(y2), None, None, 0, None, p(E_lo), emb.ld, emb.ld1, p(zbuf), zbuf.numel(),
                              s.cuda_stream)

    def timeit(fn):
        with torch.cuda.stream(s):
            for i in range(3):
                fn(i)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(8):
                    fn(i)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(8):
                g.replay()
            e1.record(s)
            e1.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 64

    print("embed_fwd_raw (ids + lookup + E_lo) alone: %.2f us" % timeit(unfused))
    print("embed_tower_fwd alone: %.2f us" % timeit(fused))
    # ---- the fused backward (dE GEMM + scatter) alone
    dpre0 = torch.randn(B, N, device=dev) * 1e-3
    dpre0_lo = ops.split_lo(dpre0)
    dy = torch.randn(B, device=dev) * 1e-3
    rows_b = []
    for i in range(8):
        fused(i)
        rows_b.append(rows.clone())
    s.synchronize()

    def fbwd(i):
        rc = lib.ctr_tower_embed_bwd(p(dpre0), p(dpre0_lo), p(W0), p(W0_lo), N, p(rows_b[i % 8]), p(E),
                                     p(S), p(dy), p(dy), emb.w1_fields, emb._offsets_host, B, F, D,
                                     p(emb.dtable), p(emb.dw1), emb.ld, emb.ld1, s.cuda_stream)
        assert rc == 0, _lib.last_error()

    dE = torch.randn(B, F * D, device=dev) * 1e-3

    def ubwd(i):
        lib.ctr_embed_bwd(p(rows_b[i % 8]), p(dE), p(E), p(emb.table), p(S), p(dy), p(dy), emb.w1_fields,
                          emb._offsets_host, B, F, D, p(emb.dtable), p(emb.dw1), emb.ld, emb.ld1,
                          s.cuda_stream)

    print("embed_bwd (scatter only) alone: %.2f us" % timeit(ubwd))
    print("tower_embed_bwd alone: %.2f us" % timeit(fbwd))
    for dbg, what in ((1, "no epilogue (TMA + MMA only)"), (2, "epilogue without REDs"),
                      (4, "all fields through the large-field path"), (6, "large-field path, no REDs"),
                      (14, "large-field path, no REDs, no FM term (no E / S loads)"),
                      (30, "large-field path, no REDs, no FM term, no match_any")):
        lib.ctr_set_option(b"eb_debug", dbg)
        print("  eb_debug=%d %s: %.2f us" % (dbg, what, timeit(fbwd)))
    lib.ctr_set_option(b"eb_debug", 0)
    tim = torch.zeros(10, dtype=torch.int64, device=dev)
    lib.ctr_embed_tower_timing(p(tim))
    accb = []
    with torch.cuda.stream(s):
        for i in range(10):
            fbwd(i)
            s.synchronize()
            t = tim.cpu().tolist()
            accb.append([(t[k + 1] - t[k]) / 1e3 for k in range(4)])
    medb = [sorted(a[k] for a in accb)[5] for k in range(4)]
    print("bwd phases of CTA (0,0), warp 2, us: setup+preloads %.2f | wait for MMAs %.2f | epilogue %.2f | "
          "CTA join %.2f" % tuple(medb))
    names = ["setup", "ids", "rid+loads issued", "A tiles written", "MMAs retired", "dump+cluster sync",
             "reduce", "stats+exit"]
    acc = []
    with torch.cuda.stream(s):
        for i in range(10):
            fused(i)
            s.synchronize()
            t = tim.cpu().tolist()
            acc.append([(t[k + 1] - t[k]) / 1e3 for k in range(7)] + [(t[8] - t[0]) / 1e3, (t[9] - t[8]) / 1e3, (t[1] - t[9]) / 1e3])
    lib.ctr_embed_tower_timing(None)
    med = [sorted(a[k] for a in acc)[5] for k in range(7)]
    sub = [sorted(a[k] for a in acc)[5] for k in range(7, 10)]
    print("ids phase split: setup (barriers, TMEM alloc) %.2f | tables staged %.2f | ids computed %.2f" % tuple(sub))
    print("phases of CTA (0,0) us: " + " | ".join("%s %.2f" % (n, m) for n, m in zip(names[1:], med)) +
          " | total %.2f" % sum(med))


if __name__ == "__main__":
    main()

"""Micro-benchmark (GPU): the row optimiser's variants on the R-full Criteo table, each as
[kernel, schedule tick] x 8 id batches captured in a CUDA graph (the claim tag comes from the device
schedule, so the graph can be replayed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from recsys_b200 import _lib, ops  # noqa: E402
from recsys_b200 import feature_column as fc  # noqa: E402
from recsys_b200.fm import fm  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()


def main():
    B = int(os.environ.get("B", 4096))
    _, cols = fm.build_feature_columns(16, full_cardinality=True)
    lay = fc.layout(cols)
    emb = ops.FieldEmbedding(lay, dev, with_w1=True, w1_fields=(1 << lay.F) - 1, adam_mode="lazy")
    F, D = lay.F, 16
    offs = torch.tensor(lay.offsets, device=dev)
    nr = (offs[1:] - offs[:-1])
    g = torch.Generator(device=dev).manual_seed(0)
    rows = [(torch.minimum((torch.rand(B, F, device=dev, generator=g) * nr.float()).long(), nr - 1)
             + offs[:-1]).to(torch.int32).contiguous() for _ in range(8)]
    st = ops.TFAdamState(lr=1e-3, device=dev)
    p = ops._p
    s = torch.cuda.Stream()

    def flat(r):
        lib.ctr_adam_rows(p(r), B * F, D, p(emb.table), p(emb._m), p(emb._v), p(emb.dtable), p(emb.w1),
                          p(emb._m1), p(emb._v1), p(emb.dw1), p(emb._claim), 0, 1e-3, 0.9, 0.999, 1e-8,
                          st.state_ptr, emb.ld, emb.ld1, emb.ldc, s.cuda_stream)

    def bf(r):
        lib.ctr_adam_rows_bf(p(r), B, F, D, p(emb.table), p(emb._m), p(emb._v), p(emb.dtable),
                             p(emb.w1), p(emb._m1), p(emb._v1), p(emb.dw1), p(emb._claim), 0, 1e-3,
                             0.9, 0.999, 1e-8, st.state_ptr, emb.ld, emb.ld1, emb.ldc, 0, s.cuda_stream)

    def timeit(fn):
        with torch.cuda.stream(s):
            for r in rows[:2]:
                fn(r)
                lib.ctr_adam_tick(st.state_ptr, 1e-3, 0.9, 0.999, s.cuda_stream)
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for r in rows:
                    fn(r)
                    lib.ctr_adam_tick(st.state_ptr, 1e-3, 0.9, 0.999, s.cuda_stream)
            gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(4):
                gr.replay()
            e1.record(s)
            e1.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 32

    def tick_only(r):
        pass
    t_tick = timeit(tick_only)
    print("tick alone %.2f us" % t_tick)
    print("flat (adam_rows_kernel<16,2>): %.2f us" % (timeit(flat) - t_tick))
    for u in (1, 2):
        lib.ctr_set_option(b"adam_rows_inflight", u)
        print("bf U=%d: %.2f us" % (u, timeit(bf) - t_tick))
    lib.ctr_set_option(b"adam_rows_inflight", 1)

if __name__ == "__main__":
    main()

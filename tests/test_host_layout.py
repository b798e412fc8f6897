"""CPU: host-side memory layouts that the kernels rely on - the row-record views of
ops.FieldEmbedding (pointer offsets and strides handed to the C ABI) and the one-blob batch
layout of estimator.GraphedTrainStep.  No compute call is made."""
import torch


def _layout(D=16):
    from recsys_b200 import feature_column as fc
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("a", 50), D),
            fc.embedding_column(fc.categorical_column_with_hash_bucket("b", 7), D)]
    return fc.layout(cols)


def test_row_record_views_alias_one_array_with_one_stride():
    from recsys_b200 import ops
    for D in (8, 16, 32):
        lay = _layout(D)
        emb = ops.FieldEmbedding(lay, torch.device("cpu"), with_w1=True, w1_fields=0b11, record=True)
        S = 4 * D + 8
        assert emb.rec.shape == (lay.total_rows, S) and emb.ld == emb.ld1 == emb.ldc == S
        base = emb.rec.data_ptr()
        # theta | m | v | g | theta1 m1 v1 g1 | claim   (include/ctr_b200.h, "Row strides")
        for view, off in ((emb.table, 0), (emb._m, D), (emb._v, 2 * D), (emb.dtable, 3 * D),
                          (emb.w1, 4 * D), (emb._m1, 4 * D + 1), (emb._v1, 4 * D + 2),
                          (emb.dw1, 4 * D + 3), (emb._claim, 4 * D + 4)):
            assert view.data_ptr() == base + 4 * off and view.stride(0) == S
        assert emb._claim.dtype == torch.int32
        assert (S * 4) % 16 == 0                       # every record starts 16-byte aligned
        # the table is initialised, everything else in the record starts at zero
        assert float(emb.table.abs().sum()) > 0 and float(emb.w1.abs().sum()) > 0
        for z in (emb._m, emb._v, emb.dtable, emb._m1, emb._v1, emb.dw1):
            assert float(z.abs().sum()) == 0.0
        t = torch.randn(lay.total_rows, D)
        emb.load(t, torch.arange(lay.total_rows, dtype=torch.float32))
        assert torch.equal(emb.rec[:, :D], t) and torch.equal(emb.rec[:, 4 * D], emb.w1)
        assert float(emb.rec[:, D:4 * D].abs().sum()) == 0.0     # load() touches theta / theta1 only


def test_planar_layout_is_the_stride_default():
    from recsys_b200 import ops
    lay = _layout()
    emb = ops.FieldEmbedding(lay, torch.device("cpu"), with_w1=True, w1_fields=0b11, record=False)
    assert emb.rec is None and (emb.ld, emb.ld1, emb.ldc) == (16, 1, 1)
    assert emb.table.is_contiguous() and emb.dtable.is_contiguous() and emb.w1.is_contiguous()
    # the dense (exact_tf) optimiser streams whole arrays: it must get planar ones
    assert not ops.FieldEmbedding(lay, torch.device("cpu"), adam_mode="exact_tf").record


def test_graphed_step_blob_layout():
    from recsys_b200.estimator import GraphedTrainStep
    ex = {"cont": torch.zeros(100, 13), "cat": torch.zeros(100, 26, dtype=torch.int64),
          "empty": torch.zeros(100, 0), "__labels__": torch.zeros(100, 1)}
    blob, views = GraphedTrainStep._make_blob(ex, torch.device("cpu"))
    assert blob.dtype == torch.uint8
    end = 0
    for k, t in ex.items():
        v = views[k]
        assert v.shape == t.shape and v.dtype == t.dtype
        if t.numel() == 0:
            continue
        off = v.data_ptr() - blob.data_ptr()
        assert off % 256 == 0 and off >= end          # slots in order, 256-byte aligned
        end = off + t.numel() * t.element_size()
    assert end <= blob.numel()
    blob.zero_()
    views["cat"].fill_(7)
    views["cont"].fill_(1.5)                          # neighbours do not overlap
    assert int(views["cat"].min()) == 7 and float(views["__labels__"].abs().sum()) == 0.0


def test_estimator_train_honours_steps_and_max_steps():
    """tf.estimator.Estimator.train(steps=, max_steps=): `steps` more steps, or up to the global step
    `max_steps` (a call at or past it trains nothing).  Host logic only: a stand-in model_fn."""
    from recsys_b200.estimator import Estimator, EstimatorSpec, ModeKeys, VariableStore

    store = VariableStore()

    def model_fn(features, labels, mode, params):
        def train_op():
            params["variable_store"].global_step += 1
        return EstimatorSpec(mode=mode, predictions={}, loss=torch.tensor(0.5), train_op=train_op)

    def input_fn():
        while True:
            yield {}, None

    est = Estimator(model_fn, params={"variable_store": store})
    assert est.train(input_fn, steps=3) == 0.5 and store.global_step == 3
    est.train(input_fn, max_steps=10)
    assert store.global_step == 10
    assert est.train(input_fn, max_steps=10) is None and store.global_step == 10
    est.train(input_fn, steps=2, max_steps=100)
    assert store.global_step == 12

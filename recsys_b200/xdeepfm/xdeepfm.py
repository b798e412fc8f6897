"""Drop-in for the hot path of reference ``xdeepfm/xdeepfm.py`` (``model_fn`` :123-233):
linear + CIN + DNN.  The CIN contraction (:145-169) runs on the tcgen05 tensor cores
(``params['cin_precision']``: 'tf32x3' default = fp32-grade, 'tf32' fastest, 'fp32' =
CUDA-core exact mode); the [B, D, 39*Hk] outer product is never materialised."""
from .. import criteo_schema as _schema
from .. import data as _data
from .._core import XDeepFMModel
from ..estimator import store_of

cont_feature = _schema.cont_feature
cat_feature = _schema.cat_feature
feature_description = _schema.feature_description


def build_feature_columns(embedding_size, full_cardinality=False):
    """xdeepfm/xdeepfm.py:44-94: linear = 13 log-numerics (:82) + 26 indicators (:91)."""
    return _schema.build_columns(embedding_size, linear="numeric+indicator",
                                 full_cardinality=full_cardinality)


def input_fn(filenames, batch_size, num_epochs=-1, need_shuffle=False):
    """xdeepfm/xdeepfm.py:103-120 (shuffle buffer 100 batches)."""
    return _data.criteo_input_fn(filenames, batch_size, num_epochs, need_shuffle, 100)


def model_fn(features, labels, mode, params):
    """xdeepfm/xdeepfm.py:123-233.  params: + deep_layers, cross_layers (comma strings, :269-277)."""
    store = store_of(params)
    model = store.get("xdeepfm", lambda: XDeepFMModel(params))
    model.store = store
    return model.spec(features, labels, mode)

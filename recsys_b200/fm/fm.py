"""Drop-in for the hot path of reference ``fm/fm.py``: same names, same argument
meaning (``build_feature_columns`` fm/fm.py:47, ``input_fn`` :106, ``model_fn`` :115).
FM = ReLU(first order over one-hots) + 0.5*sum_d[(sum_f v)^2 - sum_f v^2] -> dense(1)."""
from .. import criteo_schema as _schema
from .. import data as _data
from .._core import FMModel
from ..estimator import store_of

cont_feature = _schema.cont_feature
cat_feature = _schema.cat_feature
feature_description = _schema.feature_description


def build_feature_columns(embedding_size, full_cardinality=False):
    """fm/fm.py:47-97: linear = indicator columns of all 39 fields (:83,:94)."""
    return _schema.build_columns(embedding_size, linear="indicator_all",
                                 full_cardinality=full_cardinality)


def input_fn(filenames, batch_size, num_epochs=-1, need_shuffle=False):
    """fm/fm.py:106-112 (shuffle buffer 1000 batches)."""
    return _data.criteo_input_fn(filenames, batch_size, num_epochs, need_shuffle, 1000)


def model_fn(features, labels, mode, params):
    """fm/fm.py:115-170.  params: linear_feature_columns, embedding_feature_columns,
    embedding_size, learning_rate, dropout (:196-202)."""
    store = store_of(params)
    model = store.get("fm", lambda: FMModel(params))
    model.store = store
    return model.spec(features, labels, mode)

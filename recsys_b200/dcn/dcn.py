"""Drop-in for the hot path of reference ``dcn/dcn.py`` (``model_fn`` :117-190):
4 cross layers xl <- (xl.w) x0 + xl + b on the 624-wide embedding concat (:132-142),
fused into the gather kernel, + a BN/dropout tower; logit = dense([tower, xl])."""
from .. import criteo_schema as _schema
from .. import data as _data
from .._core import DCNModel
from ..estimator import store_of

cont_feature = _schema.cont_feature
cat_feature = _schema.cat_feature
feature_description = _schema.feature_description


def build_feature_columns(embedding_size, full_cardinality=False):
    """dcn/dcn.py:49-99: linear = the 13 log-numerics only (:86; unused by model_fn)."""
    return _schema.build_columns(embedding_size, linear="numeric",
                                 full_cardinality=full_cardinality)


def input_fn(filenames, batch_size, num_epochs=-1, need_shuffle=False):
    """dcn/dcn.py:108-114."""
    return _data.criteo_input_fn(filenames, batch_size, num_epochs, need_shuffle, 1000)


def model_fn(features, labels, mode, params):
    """dcn/dcn.py:117-190.  The reference reads the layer count from the global
    ``FLAGS.cross_layers`` (:24,:134); here it is ``params['cross_layers']`` (default 4)."""
    store = store_of(params)
    model = store.get("dcn", lambda: DCNModel(params))
    model.store = store
    return model.spec(features, labels, mode)

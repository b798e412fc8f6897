"""Drop-in for the hot path of reference ``deepfm/deepfm.py`` (``model_fn`` :73-150).
The checked-in reference is a 2-field variant (``build_model_columns`` :37-51) but its
model_fn is field-count agnostic; BASELINE's "DeepFM Criteo 39-field emb16" feeds it
the Criteo columns of fm/fm.py:47-97 - both builders are provided."""
from .. import criteo_schema as _schema
from .. import data as _data
from .. import feature_column as _fc
from .._core import DeepFMModel
from ..estimator import store_of

feature_description = {"label": ("int64", (), None), "u_id": ("int64", 1, None),
                       "i_id": ("int64", 1, None)}            # deepfm/deepfm.py:28-33
categoryFeatureNa = "####"


def build_model_columns(embedding_size):
    """deepfm/deepfm.py:37-51: u_id (500000 buckets) and i_id (100000), int64 keys hashed
    as their decimal strings [TF-sem]."""
    linear_feature_columns = []
    embedding_feature_columns = []
    u_id = _fc.categorical_column_with_hash_bucket("u_id", 500000, dtype="int64")
    linear_feature_columns.append(_fc.indicator_column(u_id))
    embedding_feature_columns.append(_fc.embedding_column(u_id, embedding_size))
    i_id = _fc.categorical_column_with_hash_bucket("i_id", 100000, dtype="int64")
    linear_feature_columns.append(_fc.indicator_column(i_id))
    embedding_feature_columns.append(_fc.embedding_column(i_id, embedding_size))
    return linear_feature_columns, embedding_feature_columns


def build_feature_columns(embedding_size, full_cardinality=False):
    """The Criteo-39 columns (fm/fm.py:47-97) for BASELINE configs 2 and 5."""
    return _schema.build_columns(embedding_size, linear="indicator_all",
                                 full_cardinality=full_cardinality)


def input_fn(filenames, batch_size=32, num_epochs=-1, need_shuffle=False):
    """deepfm/deepfm.py:60-70 applied to Criteo-shaped records (shuffle buffer 100)."""
    return _data.criteo_input_fn(filenames, batch_size, num_epochs, need_shuffle, 100)


def model_fn(features, labels, mode, params):
    """deepfm/deepfm.py:73-150.  params: + deep_layers (comma string, :172-179)."""
    store = store_of(params)
    model = store.get("deepfm", lambda: DeepFMModel(params))
    model.store = store
    return model.spec(features, labels, mode)

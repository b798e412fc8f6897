"""Drop-in for the hot path of reference ``din/din.py`` (``model_fn`` :83-180): target
item / category embeddings, two history sequences through the activation unit
(``_attention`` :103-125, fused kernel, no softmax, mask id > 0), MLP 100-50-20-1,
logit + item bias."""
import torch

from .._core import DINModel
from ..estimator import store_of

# din/din.py:43-50
feature_description = {
    "label": ("int64", (), None),
    "i_id": ("int64", (), None),
    "i_cate": ("int64", (), None),
    "u_iid_seq": ("int64", "varlen", None),
    "u_icat_seq": ("int64", "varlen", None),
}


def input_fn(filenames, batch_size, num_epochs=-1, need_shuffle=False):
    """din/din.py:63-80: VarLen sequences densified per record, ``.batch()`` (not
    padded_batch), so every record of a batch must carry the same history length.
    Decoded by the library's multi-threaded host decoder (data.din_input_fn)."""
    from ..data import din_input_fn
    return din_input_fn(filenames, batch_size, num_epochs, need_shuffle)


def model_fn(features, labels, mode, params):
    """din/din.py:83-180.  Only embedding_size, learning_rate and dropout are read from
    params, as in the reference (:217-226 vs :88-90,:118,:172)."""
    store = store_of(params)
    model = store.get("din", lambda: DINModel(params))
    model.store = store
    return model.spec(features, labels, mode)

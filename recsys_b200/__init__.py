"""recsys_b200 - B200-native (sm_100a) embedding + feature-interaction hot path of
wangruichens/recsys behind the reference's own entry points:

    recsys_b200.fm.fm / deepfm.deepfm / xdeepfm.xdeepfm / dcn.dcn / din.din
        build_feature_columns(embedding_size), input_fn(...), model_fn(features, labels, mode, params)

Compute lives in libctr_b200.so (include/ctr_b200.h), loaded with ctypes; there is
no CPU fallback.  See DESIGN.md and INTEGRATION.md.
"""
from . import estimator, feature_column  # noqa: F401

__version__ = "0.1.0"

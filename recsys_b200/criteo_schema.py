"""The Criteo 39-field schema shared by fm / deepfm / xdeepfm / dcn:
``feature_description`` (fm/fm.py:39-44) and the column builder behind each
model's ``build_feature_columns`` (fm/fm.py:47-97, xdeepfm/xdeepfm.py:44-94,
dcn/dcn.py:49-99 - identical except for what goes into the linear list)."""
from __future__ import annotations

from . import feature_column as fc

cont_feature = ["_c{0}".format(i) for i in range(0, 14)]      # _c0 is the label
cat_feature = ["_c{0}".format(i) for i in range(14, 40)]

# fm/fm.py:43-44: 14 float scalars, 26 strings defaulting to 'NULL'
feature_description = {k: ("float32", 1, None) for k in cont_feature}
feature_description.update({k: ("string", 1, "NULL") for k in cat_feature})

# fm/fm.py:54-67
buckets_cont = [
    [0.0, 1.0, 2.0, 3.0, 5.0, 12.0],
    [0.0, 1.0, 2.0, 4.0, 10.0, 28.0, 76.0, 301.0],
    [1.0, 2.0, 3.0, 5.0, 7.0, 10.0, 16.0, 24.0, 54.0],
    [1.0, 2.0, 3.0, 5.0, 6.0, 9.0, 13.0, 20.0],
    [20.0, 155.0, 1087.0, 1612.0, 2936.0, 5064.0, 8622.0, 16966.0, 39157.0],
    [3.0, 7.0, 13.0, 24.0, 36.0, 53.0, 85.0, 154.0, 411.0],
    [0.0, 1.0, 2.0, 4.0, 6.0, 10.0, 17.0, 43.0],
    [1.0, 2.0, 4.0, 6.0, 8.0, 12.0, 17.0, 25.0, 37.0],
    [4.0, 8.0, 16.0, 28.0, 41.0, 63.0, 109.0, 147.0, 321.0],
    [0.0, 1.0, 2.0],
    [0.0, 1.0, 2.0, 3.0, 4.0, 8.0],
    [0.0, 1.0, 2.0],
    [1.0, 2.0, 3.0, 5.0, 7.0, 10.0, 14.0, 22.0],
]
# fm/fm.py:72-73 (effective) and :69-70 (true cardinalities, overwritten in the reference)
buckets_cat = [1460, 583, 100000, 100000, 305, 23, 12517, 633, 3, 93145, 5683, 100000, 3194, 27,
               14992, 100000, 10, 5652, 2172, 3, 100000, 17, 15, 100000, 104, 100000]
buckets_cat_full = [1460, 583, 10131226, 2202607, 305, 23, 12517, 633, 3, 93145, 5683, 8351592,
                    3194, 27, 14992, 5461305, 10, 5652, 2172, 3, 7046546, 17, 15, 286180, 104,
                    142571]


def build_columns(embedding_size, linear="indicator_all", full_cardinality=False, hash_buckets=None):
    """linear: 'indicator_all' (fm.py:83,94), 'numeric+indicator' (xdeepfm.py:82,91),
    'numeric' (dcn.py:86; unused by its model_fn).  Unlike the reference this does not
    mutate the module-level ``cont_feature`` list (fm/fm.py:48 makes a second call raise)."""
    linear_feature_columns = []
    embedding_feature_columns = []
    cats = hash_buckets or (buckets_cat_full if full_cardinality else buckets_cat)
    for i, j in zip(cont_feature[1:], buckets_cont):
        off = 4.0 if i == "_c2" else 1.0                                   # fm/fm.py:76-78
        f_num = fc.numeric_column(i, log_offset=off)
        f_bucket = fc.bucketized_column(f_num, j)
        f_embedding = fc.embedding_column(f_bucket, embedding_size)
        if linear == "indicator_all":
            linear_feature_columns.append(fc.indicator_column(f_bucket))
        else:
            linear_feature_columns.append(f_num)
        embedding_feature_columns.append(f_embedding)
    for i, j in zip(cat_feature, cats):
        f_cat = fc.categorical_column_with_hash_bucket(key=i, hash_bucket_size=j)
        f_ind = fc.indicator_column(f_cat)
        f_embedding = fc.embedding_column(f_cat, embedding_size)
        if linear != "numeric":
            linear_feature_columns.append(f_ind)
        embedding_feature_columns.append(f_embedding)
    return linear_feature_columns, embedding_feature_columns

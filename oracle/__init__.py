"""CPU oracle for the CTR embedding + feature-interaction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``recsys_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline - never as the product path.

PARITY UNPINNED.  The reference (wangruichens/recsys @ d38b94c) holds no tests,
no golden vectors and no fixtures for this path, and all of its arithmetic
lives in TensorFlow 1.x (+ tensorflow_estimator), which is neither vendored in
/root/reference, nor version-pinned there, nor installable in this image.  The
oracle therefore restates the five ``model_fn`` graphs op for op from the
reference's own source (file:line cited on every function) plus the published
TF 1.13/1.14 semantics of the ops they call.  The only external pins are
(1) the TF-documented ``to_hash_bucket_fast`` example for the FarmHash
Fingerprint64 port and (2) the real TFRecord shard ``xdeepfm/part-r-00000``
(first 256 records frozen under ``tests/golden/``).

Layout
  criteo.py    schema, bucket boundaries, hash sizes, field order, id pipeline
  farmhash.py  FarmHash Fingerprint64 (farmhashna::Hash64) in pure Python ints
  tfsem.py     TF-1.x op semantics: initialisers, BN, dropout, BCE, AUC, Adam
  models.py    fm / deepfm / xdeepfm / dcn / din forward (torch CPU, fp64|fp32)
  tfrecord.py  TFRecord framing + tf.train.Example wire-format reader
"""

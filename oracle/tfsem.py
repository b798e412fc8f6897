"""TensorFlow-1.x op semantics the reference's model_fns rely on (oracle).

Test infrastructure - see oracle/__init__.py.  Every function names the TF op
it restates ([TF-sem]: TF 1.13/1.14 behaviour; TF itself is not vendored under
/root/reference and not installable here, so these are unpinned restatements)
and the reference call site that uses it.
"""
from __future__ import annotations

import math

import numpy as np
import torch

BN_EPS = 1e-3        # tf.layers.batch_normalization default epsilon (deepfm/deepfm.py:106)
BN_MOMENTUM = 0.99   # default momentum; moving stats are never updated by the reference
                     # because UPDATE_OPS are not attached to train_op (deepfm/deepfm.py:142-143)


# ----------------------------------------------------------------- initialisers
def truncated_normal(shape, stddev, gen: torch.Generator, dtype=torch.float64):
    """tf.truncated_normal_initializer: N(0, stddev) resampled outside 2 sigma.
    embedding_column default: stddev = 1/sqrt(dimension) (fm/fm.py:80,90)."""
    out = torch.randn(shape, generator=gen, dtype=torch.float64)
    bad = out.abs() > 2
    while bad.any():
        out[bad] = torch.randn(int(bad.sum()), generator=gen, dtype=torch.float64)
        bad = out.abs() > 2
    return (out * stddev).to(dtype)


def _fans(shape):
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def glorot_uniform(shape, gen, dtype=torch.float64):
    """tf.layers.dense / tf.get_variable default kernel initialiser
    (deepfm/deepfm.py:91,104; xdeepfm/xdeepfm.py:154-156)."""
    fi, fo = _fans(tuple(shape))
    lim = math.sqrt(6.0 / (fi + fo))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)


def glorot_normal(shape, gen, dtype=torch.float64):
    """tf.glorot_normal_initializer (dcn/dcn.py:139-140, din/din.py:89-90):
    truncated normal, variance 2/(fan_in+fan_out)."""
    fi, fo = _fans(tuple(shape))
    std = math.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
    return truncated_normal(shape, std, gen, dtype)


# ------------------------------------------------------------------------ layers
def dense(x, w, b, relu=False):
    """tf.layers.dense: x @ kernel + bias (+ ReLU)."""
    y = x @ w + b
    return torch.relu(y) if relu else y


def batch_norm(x, gamma, beta, mean, var, training):
    """tf.layers.batch_normalization(training=...): training -> batch mean and
    *biased* batch variance; inference -> moving stats.  eps 1e-3."""
    if training:
        mu = x.mean(0)
        va = x.var(0, unbiased=False)
    else:
        mu, va = mean, var
    return (x - mu) / torch.sqrt(va + BN_EPS) * gamma + beta


def dropout(x, rate, training, mask=None):
    """tf.layers.dropout: inverted dropout, keep prob 1-rate.  ``mask`` (0/1
    keep mask) is injected for parity runs because TF's RNG stream cannot be
    reproduced."""
    if not training or rate == 0.0:
        return x
    if mask is None:
        raise ValueError("oracle dropout needs an explicit keep mask when rate > 0")
    return x * mask.to(x.dtype) / (1.0 - rate)


def sigmoid_cross_entropy_with_logits(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits (fm/fm.py:146-149), written as TF writes it
    so that autodiff agrees at logits == 0 exactly (all three ReLU branches dead happens):
        cond = x >= 0; relu = where(cond, x, 0); neg_abs = where(cond, -x, x)
        loss = relu - x*z + log1p(exp(neg_abs))            -> d/dx = sigmoid(x) - z everywhere
    (clamp/abs would give 1 - z at x == 0)."""
    cond = logits >= 0
    zeros = torch.zeros_like(logits)
    relu_logits = torch.where(cond, logits, zeros)
    neg_abs = torch.where(cond, -logits, logits)
    return relu_logits - logits * labels + torch.log1p(torch.exp(neg_abs))


# ----------------------------------------------------------------------- metrics
class StreamingAUC:
    """tf.metrics.auc(labels, pred): 200 thresholds, ROC, trapezoid (fm/fm.py:151).

    thresholds = [-1e-7, 1/199, ..., 198/199, 1+1e-7]; positive iff pred > thr;
    tpr = (tp+1e-6)/(tp+fn+1e-6), fpr = fp/(fp+tn+1e-6); fp32 accumulators.
    """

    def __init__(self, num_thresholds=200):
        n = num_thresholds
        thr = [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)]
        self.thr = np.array([0.0 - 1e-7] + thr + [1.0 + 1e-7], np.float32)
        self.tp = np.zeros(n, np.float32)
        self.fp = np.zeros(n, np.float32)
        self.tn = np.zeros(n, np.float32)
        self.fn = np.zeros(n, np.float32)

    def update(self, labels, pred):
        y = np.asarray(labels).reshape(-1).astype(bool)
        p = np.asarray(pred, np.float32).reshape(-1)
        pos = p[None, :] > self.thr[:, None]
        self.tp += (pos & y[None]).sum(1).astype(np.float32)
        self.fp += (pos & ~y[None]).sum(1).astype(np.float32)
        self.fn += (~pos & y[None]).sum(1).astype(np.float32)
        self.tn += (~pos & ~y[None]).sum(1).astype(np.float32)

    def result(self):
        eps = np.float32(1e-6)
        rec = (self.tp + eps) / (self.tp + self.fn + eps)
        fpr = self.fp / (self.fp + self.tn + eps)
        return float(np.sum((fpr[:-1] - fpr[1:]) * (rec[:-1] + rec[1:]) / np.float32(2.0)))


class StreamingAccuracy:
    """tf.metrics.accuracy(labels, tf.round(pred)) (fm/fm.py:152); tf.round is
    round-half-to-even."""

    def __init__(self):
        self.total = 0.0
        self.count = 0.0

    def update(self, labels, pred):
        y = np.asarray(labels, np.float32).reshape(-1)
        p = np.rint(np.asarray(pred, np.float32).reshape(-1))
        self.total += float((y == p).sum())
        self.count += float(y.size)

    def result(self):
        return self.total / max(self.count, 1.0)


# --------------------------------------------------------------------- optimiser
class TFAdam:
    """tf.train.AdamOptimizer(lr) (fm/fm.py:162-163): beta1 .9, beta2 .999,
    eps 1e-8 *outside* the bias-corrected sqrt:
        lr_t = lr*sqrt(1-b2^t)/(1-b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
        theta -= lr_t * m / (sqrt(v) + eps)
    Sparse (IndexedSlices) gradients are de-duplicated by segment-sum and then
    m, v and theta are updated for *every* row ([TF-sem] _apply_sparse_shared),
    i.e. exactly the dense rule applied to the scattered dense gradient.
    ``lazy=True`` is the non-reference LazyAdam variant (touched rows only).
    """

    def __init__(self, params: dict, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params: dict, grads: dict, lazy_rows: dict | None = None):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, g in grads.items():
            if g is None:
                continue
            p, m, v = params[k], self.m[k], self.v[k]
            if lazy_rows is not None and k in lazy_rows:
                r = torch.unique(lazy_rows[k])
                m[r] = self.b1 * m[r] + (1 - self.b1) * g[r]
                v[r] = self.b2 * v[r] + (1 - self.b2) * g[r] * g[r]
                p[r] -= lr_t * m[r] / (v[r].sqrt() + self.eps)
            else:
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * m / (v.sqrt() + self.eps))

"""The Estimator callback contract of the reference, mirrored for eager PyTorch.

The only stable interface of the reference's hot path is
``model_fn(features, labels, mode, params) -> EstimatorSpec`` (fm/fm.py:115,135-170
and the same block in deepfm/xdeepfm/dcn/din).  This module keeps those names:
``ModeKeys``, ``EstimatorSpec``, ``export.PredictOutput``, the two streaming
metrics, and a variable store that plays the role of the TF graph's variable
collection (variables are created by the first ``model_fn`` call and reused by
later ones).  ``Estimator`` is a thin eager train/evaluate/predict loop - the
reference's checkpointing / summaries / distribution strategy are out of scope.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import torch


class ModeKeys:
    TRAIN = "train"
    EVAL = "eval"
    PREDICT = "infer"


@dataclass
class PredictOutput:
    outputs: dict


class export:  # namespace mirror of estimator.export
    PredictOutput = PredictOutput


DEFAULT_SERVING_SIGNATURE_DEF_KEY = "serving_default"


@dataclass
class EstimatorSpec:
    mode: str
    predictions: Optional[dict] = None
    loss: Optional[torch.Tensor] = None
    train_op: Optional[Callable[[], None]] = None
    eval_metric_ops: Optional[dict] = None
    export_outputs: Optional[dict] = None


# ----------------------------------------------------------------------- metrics
class StreamingAUC:
    """tf.metrics.auc(labels, pred) (fm/fm.py:151): 200 thresholds
    [-1e-7, 1/199..198/199, 1+1e-7], positive iff pred > thr, trapezoid over
    (fpr, tpr) with tpr=(tp+1e-6)/(tp+fn+1e-6), fpr=fp/(fp+tn+1e-6); fp32 counters.
    State lives on the device of the first update."""

    def __init__(self, num_thresholds: int = 200):
        n = num_thresholds
        thr = [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)]
        self._thr_host = torch.tensor([0.0 - 1e-7] + thr + [1.0 + 1e-7], dtype=torch.float32)
        self.thr = None
        self.tp = self.fp = self.tn = self.fn = None

    def update(self, labels, pred):
        pred = pred.detach().reshape(-1).float()
        y = labels.detach().reshape(-1).to(pred.device) > 0.5
        if self.thr is None:
            self.thr = self._thr_host.to(pred.device)
            z = torch.zeros_like(self.thr)
            self.tp, self.fp, self.tn, self.fn = z.clone(), z.clone(), z.clone(), z.clone()
        pos = pred[None, :] > self.thr[:, None]
        yy = y[None, :]
        self.tp += (pos & yy).sum(1).float()
        self.fp += (pos & ~yy).sum(1).float()
        self.fn += (~pos & yy).sum(1).float()
        self.tn += (~pos & ~yy).sum(1).float()
        return self

    def result(self) -> float:
        eps = 1e-6
        rec = (self.tp + eps) / (self.tp + self.fn + eps)
        fpr = self.fp / (self.fp + self.tn + eps)
        return float(((fpr[:-1] - fpr[1:]) * (rec[:-1] + rec[1:]) / 2.0).sum())


class StreamingAccuracy:
    """tf.metrics.accuracy(labels, tf.round(pred)) (fm/fm.py:152)."""

    def __init__(self):
        self.total = 0.0
        self.count = 0.0

    def update(self, labels, pred):
        p = torch.round(pred.detach().reshape(-1).float())
        y = labels.detach().reshape(-1).to(p.device).float()
        self.total += float((p == y).sum())
        self.count += float(y.numel())
        return self

    def result(self) -> float:
        return self.total / max(self.count, 1.0)


# ---------------------------------------------------------------- variable store
class VariableStore:
    """Stands in for the TF graph's variable collection: ``model_fn`` asks for its
    model object by scope name; the first call builds it, later calls reuse it."""

    def __init__(self):
        self._objs: Dict[str, object] = {}
        self.global_step = 0

    def get(self, scope: str, factory: Callable[[], object]):
        if scope not in self._objs:
            self._objs[scope] = factory()
        return self._objs[scope]

    def reset(self):
        self._objs.clear()
        self.global_step = 0


_DEFAULT_STORE = VariableStore()


def default_store() -> VariableStore:
    return _DEFAULT_STORE


def store_of(params: dict) -> VariableStore:
    """``params['variable_store']`` when the caller wants isolation, else the
    process-wide default (the reference has one graph per Estimator)."""
    return params.get("variable_store") or _DEFAULT_STORE


# --------------------------------------------------------------------- Estimator
class Estimator:
    """Eager stand-in for ``tf.estimator.Estimator(model_fn, model_dir, params, config)``
    limited to the calls the reference drivers make (fm/fm.py:204-224)."""

    def __init__(self, model_fn, model_dir=None, params=None, config=None):
        self.model_fn = model_fn
        self.params = dict(params or {})
        self.params.setdefault("variable_store", VariableStore())
        self.model_dir = model_dir
        self.config = config

    def train(self, input_fn, steps=None, max_steps=None):
        """``steps``: train that many more steps; ``max_steps``: stop once the global step has
        reached it (a call at or past it trains nothing) - tf.estimator.Estimator.train."""
        store = store_of(self.params)
        n = 0
        last = None
        if max_steps is not None and store.global_step >= max_steps:
            return None
        for features, labels in input_fn():
            spec = self.model_fn(features, labels, ModeKeys.TRAIN, self.params)
            spec.train_op()
            last = spec.loss
            n += 1
            if steps is not None and n >= steps:
                break
            if max_steps is not None and store.global_step >= max_steps:
                break
        return None if last is None else float(last)

    def evaluate(self, input_fn, steps=None):
        auc, acc = StreamingAUC(), StreamingAccuracy()
        tot, n = 0.0, 0
        for features, labels in input_fn():
            spec = self.model_fn(features, labels, ModeKeys.EVAL, self.params)
            auc.update(labels, spec.predictions["prob"])
            acc.update(labels, spec.predictions["prob"])
            tot += float(spec.loss)
            n += 1
            if steps is not None and n >= steps:
                break
        return {"AUC": auc.result(), "Accuracy": acc.result(), "loss": tot / max(n, 1),
                "global_step": self.params["variable_store"].global_step}

    def predict(self, input_fn):
        for features, labels in input_fn():
            spec = self.model_fn(features, None, ModeKeys.PREDICT, self.params)
            yield from ({"prob": p} for p in spec.predictions["prob"].detach().cpu())


# ------------------------------------------------------------------ graphed step
class GraphedTrainStep:
    """One whole train step - ``spec = model_fn(features, labels, TRAIN, params);
    spec.train_op()`` - captured into CUDA graphs over static device input buffers and
    replayed per batch.

    The step is captured TWICE, over two static input blobs, and the replays alternate: while
    graph k runs on the step stream, the next batch is copied straight into the other blob on a
    copy stream (pinned host tensors: H2D; batches already in HBM: one D2D), so no copy and no
    copy-to-graph dependency sits between two replays.  ``__call__(features, labels)`` returns
    the device loss tensor without synchronising; ``loss_to_host`` reads it back on a third
    stream so that the read does not sit between two replays.  The loss is copied out of the
    step's workspace on an auxiliary stream right after the forward (beside the backward)."""

    def __init__(self, model_fn, params, example_features, example_labels, warmup: int = 3,
                 double_buffer: Optional[bool] = None):
        from .ops import PackedFeatures
        dev = params.get("device") or torch.device("cuda", torch.cuda.current_device())
        self.model_fn, self.params = model_fn, params
        cur = torch.cuda.current_stream(dev)
        # capture on the caller's stream when it is already a side stream (keeps every autograd
        # node on one stream); the legacy default stream cannot be captured
        self.stream = cur if cur != torch.cuda.default_stream(dev) else \
            torch.cuda.Stream(device=dev, priority=-1)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.d2h_stream = torch.cuda.Stream(device=dev)
        self.aux_stream = torch.cuda.Stream(device=dev)
        if double_buffer is None:
            double_buffer = os.environ.get("CTR_GRAPH_DOUBLE", "1") != "0"
        self.nbuf = 2 if double_buffer else 1
        f = example_features
        self._packed = isinstance(f, PackedFeatures)
        ex = {"cont": f.cont, "cat": f.cat} if self._packed else \
            {k: torch.as_tensor(v) for k, v in f.items()}
        ex["__labels__"] = example_labels
        # static inputs (what graph k reads): ONE blob with views per buffer
        self._blobs, self._statics, self._labels, self._features = [], [], [], []
        for _ in range(self.nbuf):
            blob, static = self._make_blob(ex, dev)
            self._blobs.append(blob)
            self._labels.append(static.pop("__labels__"))
            self._statics.append(static)
            if self._packed:
                self._features.append(PackedFeatures(static["cont"], static["cat"], f.cont_keys,
                                                     f.cat_keys))
            else:       # plain dict of tensors (e.g. the DIN features)
                self._features.append(dict(static))
        # one loss slot per graph: the read-back of step s (third stream) then only has to finish
        # before step s+2 rewrites its slot, and never holds up step s+1
        self._loss = [torch.zeros((), dtype=torch.float32, device=dev) for _ in range(self.nbuf)]
        self._ready = [torch.cuda.Event() for _ in range(self.nbuf)]     # batch landed in blob k
        self._free = [torch.cuda.Event() for _ in range(self.nbuf)]      # graph k done with blob k
        self._loss_ready = torch.cuda.Event()
        self._loss_read = [torch.cuda.Event() for _ in range(self.nbuf)]
        for e in self._free + self._loss_read:
            e.record(self.stream)
        self._slot = 0          # the buffer the next batch goes to
        self._last = 0          # the buffer of the latest batch
        self._prefetch = None   # the model's id pipeline, run on the copy stream (_find_prefetcher)
        for k in range(self.nbuf):
            self._slot = k
            self._load(example_features, example_labels)
        self._slot = 0
        self.stream.wait_stream(self.copy_stream)
        with torch.cuda.stream(self.stream):
            for w in range(max(warmup, 1)):         # allocator + lazy-init warm-up, eager
                self._eager(0)
                if w == 0:
                    self._find_prefetcher()         # the model object exists after its first call
        self.stream.synchronize()
        import gc
        gc.collect()
        self.graphs = []
        for k in range(self.nbuf):
            g = torch.cuda.CUDAGraph()
            kw = {"pool": self.graphs[0].pool()} if self.graphs else {}
            with torch.cuda.graph(g, stream=self.stream, **kw):
                self._eager(k)
            self.graphs.append(g)
            self.stream.synchronize()
        # capturing ran train_op's host-side bookkeeping without running a step
        store_of(self.params).global_step -= len(self.graphs)
        self.graph = self.graphs[0]

    def _find_prefetcher(self):
        """A model with a ``prefetch_ids(features)`` method gets the id pipeline of batch s+1 run on
        the copy stream, right behind the batch's copy and beside the compute of step s: the step's
        lookup kernel then starts from ready [B, F] ids (``PackedFeatures.rows``)."""
        if not self._packed or os.environ.get("CTR_PREFETCH_IDS", "1") == "0":
            return
        cands = [m for m in store_of(self.params)._objs.values() if hasattr(m, "prefetch_ids")]
        if len(cands) != 1 or not hasattr(cands[0], "F"):
            return
        m = cands[0]
        for f in self._features:
            B = (f.cat if f.cat is not None else f.cont).shape[0]
            f.rows = torch.empty((B, m.F), dtype=torch.int32, device=self._blobs[0].device)
        if all(m.prefetch_ids(f) for f in self._features):
            self._prefetch = m.prefetch_ids
        else:
            for f in self._features:
                f.rows = None

    # buffer 0's views under the old names (tests, callers that fill the static inputs themselves)
    @property
    def features(self):
        return self._features[0]

    @property
    def labels(self):
        return self._labels[0]

    @property
    def _blob(self):
        return self._blobs[0]

    @property
    def _static(self):
        return self._statics[0]

    @staticmethod
    def _make_blob(example: dict, dev):
        offs, n = {}, 0
        for k, t in example.items():
            offs[k] = n
            n += (t.numel() * t.element_size() + 255) // 256 * 256
        blob = torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
        views = {k: blob[offs[k]:offs[k] + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
                 for k, t in example.items()}
        return blob, views

    def _eager(self, k: int = 0):
        spec = self.model_fn(self._features[k], self._labels[k], ModeKeys.TRAIN, self.params)
        # the loss leaves the step's workspace on the auxiliary stream, beside the backward
        main, aux = torch.cuda.current_stream(), self.aux_stream
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            self._loss[k].copy_(spec.loss)
        spec.train_op()
        main.wait_stream(aux)

    def _sources(self, features, labels):
        src = {"cont": features.cont, "cat": features.cat} if self._packed else \
            {k: torch.as_tensor(features[k]) for k in self._statics[0]}
        src["__labels__"] = labels
        return src

    def _load(self, features, labels):
        """Copy a batch into the next buffer on the copy stream (beside the running replay)."""
        k = self._slot
        src = self._sources(features, labels)
        dst = dict(self._statics[k])
        dst["__labels__"] = self._labels[k]
        cs = self.copy_stream
        cs.wait_event(self._free[k])                 # graph k's last replay has consumed blob k
        with torch.cuda.stream(cs):
            for key, t in src.items():
                dst[key].copy_(t, non_blocking=True)
            if self._prefetch is not None:
                self._prefetch(self._features[k])
            self._ready[k].record(cs)
        return k

    def _replay(self, k):
        self.stream.wait_event(self._ready[k])
        with torch.cuda.stream(self.stream):
            self.graphs[k].replay()
            self._free[k].record(self.stream)
        self._last = k
        self._slot = (k + 1) % self.nbuf
        store_of(self.params).global_step += 1       # the captured train_op's host-side bookkeeping
        return self.loss

    @property
    def loss(self):
        """Device loss of the latest step."""
        return self._loss[self._last]

    def wait_loss_slot(self):
        """Order the next replay behind the read-back of the loss slot it is going to rewrite."""
        self.stream.wait_event(self._loss_read[self._slot])

    def __call__(self, features, labels):
        k = self._load(features, labels)
        self.stream.wait_event(self._loss_read[k])   # slot k's previous loss has been read back
        return self._replay(k)

    def to_device_batch(self, features, labels) -> torch.Tensor:
        """Lay a batch out on the device exactly like a static input blob (for batches that are
        kept resident in HBM); ``run_device_batch`` then needs one D2D copy per step."""
        src = self._sources(features, labels)
        ex = dict(self._statics[0])
        ex["__labels__"] = self._labels[0]
        blob, views = self._make_blob(ex, self._blobs[0].device)
        with torch.cuda.stream(self.stream):
            for k, t in src.items():
                views[k].copy_(t, non_blocking=True)
        self.stream.synchronize()
        return blob

    def pin_batch(self, features, labels) -> torch.Tensor:
        """The batch as ONE pinned host blob in the static buffers' layout (what an input pipeline
        that decodes straight into pinned batch buffers hands over): ``run_device_batch`` then
        moves it with a single host-to-device copy."""
        src = self._sources(features, labels)
        ex = dict(self._statics[0])
        ex["__labels__"] = self._labels[0]
        blob, views = self._make_blob(ex, torch.device("cpu"))
        blob = blob.pin_memory()
        offs, n = {}, 0
        for k, t in ex.items():
            offs[k] = n
            n += (t.numel() * t.element_size() + 255) // 256 * 256
        for k, t in src.items():
            e = ex[k]
            blob[offs[k]:offs[k] + e.numel() * e.element_size()].view(e.dtype).view(e.shape).copy_(
                torch.as_tensor(t).to(e.dtype).reshape(e.shape))
        return blob

    def run_device_batch(self, blob: torch.Tensor):
        """One copy of a blob laid out like the static buffers (device resident: D2D; pinned host:
        H2D) into the next buffer on the copy stream, then the replay."""
        k = self._slot
        cs = self.copy_stream
        cs.wait_event(self._free[k])
        with torch.cuda.stream(cs):
            self._blobs[k].copy_(blob, non_blocking=True)
            if self._prefetch is not None:
                self._prefetch(self._features[k])
            self._ready[k].record(cs)
        return self._replay(k)

    def replay_resident(self):
        """Replay on whatever is in the latest static buffer (inputs already in HBM)."""
        with torch.cuda.stream(self.stream):
            self.graphs[self._last].replay()
        return self.loss

    def loss_to_host(self, pinned_slot: torch.Tensor):
        self._loss_ready.record(self.stream)
        self.d2h_stream.wait_event(self._loss_ready)
        with torch.cuda.stream(self.d2h_stream):
            pinned_slot.copy_(self._loss[self._last], non_blocking=True)
            self._loss_read[self._last].record(self.d2h_stream)

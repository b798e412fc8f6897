"""Model objects behind the five reference ``model_fn`` entry points.

Each class holds the variables the reference's graph would create, runs the
forward through the fused sm_100a kernels (ops.py) and the small dense towers
through torch (cuBLAS; SURVEY 2.3 K7: "stays torch"), and exposes
``spec(features, labels, mode)`` returning the EstimatorSpec the reference's
model_fn returns.  Parameter names follow DESIGN.md "parameter names" (shared by
convention with the oracle so tests can load identical weights).
"""
from __future__ import annotations

import os

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as Fn

from . import feature_column as fc
from . import ops
from .estimator import (DEFAULT_SERVING_SIGNATURE_DEF_KEY, EstimatorSpec, ModeKeys, PredictOutput,
                        StreamingAccuracy, StreamingAUC, VariableStore)

BN_EPS = 1e-3   # tf.layers.batch_normalization default (deepfm/deepfm.py:106)


def _glorot_uniform(shape, gen):
    if len(shape) == 1:
        fi = fo = shape[0]
    else:
        fi, fo = shape[0], shape[1]
    lim = math.sqrt(6.0 / (fi + fo))
    return (torch.rand(shape, generator=gen) * 2 - 1) * lim


def _glorot_normal(shape, gen):
    if len(shape) == 1:
        fi = fo = shape[0]
    else:
        fi, fo = shape[0], shape[1]
    std = math.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t


def _device(params):
    dev = params.get("device")
    if dev is None:
        if not torch.cuda.is_available():
            raise RuntimeError("recsys_b200 needs a CUDA device (B200, sm_100); there is no CPU "
                               "path - the oracle under oracle/ is test infrastructure only")
        dev = torch.device("cuda", torch.cuda.current_device())
    return torch.device(dev)


def bce_with_logits_mean(logits, labels):
    """tf.reduce_mean(tf.nn.sigmoid_cross_entropy_with_logits) (fm/fm.py:146-149)."""
    z = labels.to(logits.device, torch.float32).reshape(logits.shape)
    return Fn.binary_cross_entropy_with_logits(logits, z)


class _ModelBase:
    name = "base"

    def __init__(self, params: dict):
        self.params = params
        self.device = _device(params)
        self.dropout = float(params.get("dropout", 0.0))
        self.adam = ops.TFAdamState(lr=float(params.get("learning_rate", 1e-3)), device=self.device)
        self.store: Optional[VariableStore] = None
        self.last = {}

    # towers ---------------------------------------------------------------------
    def _tower_shapes(self, shapes, prefix, sizes, bn):
        for l, (i, o) in enumerate(zip(sizes[:-1], sizes[1:])):
            shapes[f"{prefix}.{l}.w"] = (i, o)
            shapes[f"{prefix}.{l}.b"] = (o,)
            if bn:
                for s in ("gamma", "beta", "mean", "var"):
                    shapes[f"{prefix}.{l}.bn.{s}"] = (o,)

    def _init_dense(self, seed):
        g = torch.Generator().manual_seed(seed)
        P = self.dense
        with torch.no_grad():
            for n in P.names:
                v = P[n]
                if n.endswith(".w"):
                    v.copy_(self._init_weight(n, tuple(v.shape), g))
                elif n.endswith((".bn.gamma", ".bn.var")):
                    v.fill_(1.0)
                else:
                    v.zero_()

    def _init_weight(self, name, shape, g):
        return _glorot_uniform(shape, g)

    def _tower(self, x, prefix, n_layers, training, bn=True):
        P = self.dense
        for l in range(n_layers):
            x = torch.relu(torch.addmm(P[f"{prefix}.{l}.b"], x, P[f"{prefix}.{l}.w"]))
            if bn:
                if training:   # batch mean / biased variance; moving stats untouched
                    x = Fn.batch_norm(x, None, None, P[f"{prefix}.{l}.bn.gamma"],
                                      P[f"{prefix}.{l}.bn.beta"], True, 0.0, BN_EPS)
                else:
                    x = Fn.batch_norm(x, P[f"{prefix}.{l}.bn.mean"], P[f"{prefix}.{l}.bn.var"],
                                      P[f"{prefix}.{l}.bn.gamma"], P[f"{prefix}.{l}.bn.beta"],
                                      False, 0.0, BN_EPS)
            if training and self.dropout > 0.0:
                x = Fn.dropout(x, self.dropout, True)
        return x

    # state ----------------------------------------------------------------------
    def load_state(self, state: Dict[str, torch.Tensor]):
        """Load oracle-named parameters (tests, checkpoints)."""
        self.dense.load(state)

    def dense_grads(self):
        tw = getattr(self, "tower", None)
        if tw is not None:
            tw.join()
        return self.dense.grads()

    # EstimatorSpec --------------------------------------------------------------
    def logits(self, features, mode):
        raise NotImplementedError

    def _apply_gradients(self):
        raise NotImplementedError

    def forward(self, features, labels, training):
        """-> (logits, prob, loss or None).  Default: ``logits()`` + torch sigmoid / BCE; the
        Criteo models with a fused loss head override it."""
        logits = self.logits(features, training)
        pred = torch.sigmoid(logits)
        loss = bce_with_logits_mean(logits, labels) if labels is not None else None
        return logits, pred, loss

    # A sticky device word collects what the kernels cannot raise themselves: bit 0 = a categorical
    # id outside [0, n_rows) (wrapped by modulo), bit 1 = a DIN history id outside its table (read
    # as padding), bit 2 = a sharded exchange slab overflowed (lookups dropped).  It is read - one
    # 4-byte D2H copy and a synchronisation - on every eval / predict call and every
    # ``status_every`` (default 256) train calls, never inside a CUDA-graph capture.
    STATUS_TEXT = {1: "categorical id outside [0, num_buckets) (wrapped by modulo; hash raw ids first)",
                   2: "DIN history id outside its table (treated as padding)",
                   4: "sharded exchange slab overflow: lookups were dropped (raise shard_slack)"}

    def status_word(self) -> Optional[torch.Tensor]:
        ids = getattr(self, "ids", None)
        return None if ids is None else ids.status

    def check_status(self):
        w = self.status_word()
        if w is None or torch.cuda.is_current_stream_capturing():
            return
        v = int(w.item())
        if v:
            w.zero_()
            raise RuntimeError("recsys_b200 %s: %s" % (self.name, "; ".join(
                t for b, t in self.STATUS_TEXT.items() if v & b)))

    def spec(self, features, labels, mode) -> EstimatorSpec:
        training = mode == ModeKeys.TRAIN
        self._training = training
        with torch.set_grad_enabled(training):
            logits, pred, loss = self.forward(features, None if mode == ModeKeys.PREDICT else labels,
                                              training)
        self._calls = getattr(self, "_calls", 0) + 1
        if not training or self._calls % int(self.params.get("status_every", 256)) == 0:
            self.check_status()
        predictions = {"prob": pred}
        export_outputs = {DEFAULT_SERVING_SIGNATURE_DEF_KEY: PredictOutput(predictions)}
        if mode == ModeKeys.PREDICT:
            return EstimatorSpec(mode=mode, predictions=predictions, export_outputs=export_outputs)
        if mode == ModeKeys.EVAL:
            ops_ = {"AUC": StreamingAUC().update(labels, pred),
                    "Accuracy": StreamingAccuracy().update(labels, pred)}
            return EstimatorSpec(mode=mode, predictions=predictions, loss=loss.detach(),
                                 eval_metric_ops=ops_)

        def train_op():
            lr_t = self.adam.next_lr_t()
            self._arm_row_optimiser(lr_t)     # scatter-add + Adam of the tables in one pass
            self.backward(loss)
            self._apply_gradients(lr_t)
            if self.store is not None:
                self.store.global_step += 1

        self.last = {"loss": loss, "logits": logits}
        return EstimatorSpec(mode=mode, predictions=predictions, loss=loss.detach(),
                             train_op=train_op)

    def backward(self, loss):
        if getattr(self, "_one", None) is None:        # root gradient, made once (no fill per step)
            self._one = torch.ones((), dtype=torch.float32, device=self.device)
        loss.backward(self._one.expand_as(loss))
        # drop the autograd graph: a live graph keeps the leaves' AccumulateGrad nodes (and the
        # stream they were created on) alive, which breaks a later CUDA-graph capture
        self.last = {k: v.detach() for k, v in self.last.items()}

    def _arm_row_optimiser(self, lr_t):
        """params['fused_row_adam'] = True runs a table's scatter-add and row Adam in one pass
        inside the backward (ctr_count_rows + ctr_embed_bwd_adam).  Default False: measured on B200
        (profiles/r02_*), on one GPU the extra counting pass costs more than the second visit it
        saves - 51 vs 35 us at batch 4096 x 39 fields, 469 vs 352 us at batch 65 536 - so the
        unfused pair stays the default.  The peer-memory sharded table always fuses: its owner
        counts the lookups for free while it gathers the rows (K2), and applies them in K5."""
        want = bool(self.params.get("fused_row_adam", False))
        for emb in (getattr(self, "emb", None), getattr(self, "emb_dnn", None)):
            if emb is None or not getattr(emb, "can_fuse", False) or self.rows is None:
                continue
            if want or getattr(emb, "p2p", False):
                emb.arm_fused(self.rows, lr_t, self.adam)

    def apply_gradients(self):
        """The unfused optimiser step (after a plain ``backward``): Adam over the accumulated
        gradients.  ``train_op`` fuses the tables' share into the backward instead."""
        lr_t = self.adam.next_lr_t()
        self._apply_gradients(lr_t)
        if self.store is not None:
            self.store.global_step += 1


class _CriteoBase(_ModelBase):
    """Shared by fm / deepfm / xdeepfm / dcn: one FieldEmbedding over the
    embedding columns + the id pipeline."""

    want_w1 = True

    def __init__(self, params):
        super().__init__(params)
        self.lay = fc.layout(params["embedding_feature_columns"])
        if int(params.get("embedding_size", self.lay.dimension)) != self.lay.dimension:
            raise ValueError("params['embedding_size'] disagrees with the embedding columns")
        self.F, self.D = self.lay.F, self.lay.dimension
        mask, self.numeric_linear = fc.first_order_fields(params.get("linear_feature_columns", []),
                                                          self.lay) if self.want_w1 else (0, [])
        seed = int(params.get("seed", 0))
        self.world = 1
        if params.get("shard_embedding"):
            # row-sharded table over the ranks of the default process group: exchanged over NVLink
            # peer memory by our own kernels (p2p.py; the default on CUDA), or through NCCL
            # all-to-alls (sharded.py; also what the gloo CPU tests drive)
            import torch.distributed as dist
            self.world = dist.get_world_size()
            exchange = params.get("shard_exchange") or \
                ("p2p" if (self.device.type == "cuda" and params.get("shard_ops") is None) else "nccl")
            if exchange == "p2p":
                from . import p2p
                if params.get("embedding_adam", "lazy") != "lazy":
                    raise ValueError("shard_exchange='p2p' implements the lazy row optimiser only")
                self.emb = p2p.P2PShardedEmbedding(self.lay, self.device, with_w1=self.want_w1,
                                                   w1_fields=mask, seed=seed)
            else:
                from . import sharded
                self.emb = sharded.ShardedFieldEmbedding(
                    self.lay, self.device, with_w1=self.want_w1, w1_fields=mask,
                    adam_mode=params.get("embedding_adam", "lazy"), seed=seed,
                    slack=float(params.get("shard_slack", 1.5)), shard_ops=params.get("shard_ops"))
        else:
            self.emb = ops.FieldEmbedding(self.lay, self.device, with_w1=self.want_w1,
                                          w1_fields=mask,
                                          adam_mode=params.get("embedding_adam", "lazy"), seed=seed)
        self.ids = ops.IdPipeline(self.lay, self.device)
        if hasattr(self.emb, "status"):
            self.emb.status = self.ids.status       # slab overflow raises through the status word
        self.rows = None
        # fused tower + loss head kernels (tower.cu); False keeps the torch (cuBLAS) tower
        self.fused = bool(params.get("fused_tower", True))
        self._head_anchor = torch.zeros((), device=self.device, requires_grad=True)

    def load_state(self, state):
        super().load_state(state)
        self.emb.load(state.get("emb"), state.get("w1"))

    def prefetch_ids(self, features) -> bool:
        """Compute the batch's row ids into ``features.rows`` on the CURRENT stream - what
        estimator.GraphedTrainStep runs on its copy stream right after the batch has landed, so the
        id pipeline of step s+1 overlaps the compute of step s.  False when this model's step needs
        more than the ids from the id stage (the log-normalised numerics of xDeepFM)."""
        if getattr(self, "needs_logx", False) or self.world != 1 or \
                getattr(features, "rows", None) is None or not hasattr(self.emb, "lookup_features"):
            return False
        # a step whose kernels take whole SMs (the fused lookup + first layer kernel, ctr_tower_mid:
        # one CTA per SM on 128 of the 148 SMs) gets the few-CTA id kernel, so that it can never keep
        # one of those CTAs waiting; short steps (FM) get the full-width one, which is done sooner
        tw = getattr(self, "tower", None)
        bg = int(os.environ.get("CTR_IDS_BG_CTAS", "20")) if (tw is not None and tw.use_mid) else 0
        self.ids(features, out=features.rows, background=bg)
        return True

    @staticmethod
    def _batch_size(features):
        if isinstance(features, ops.PackedFeatures):
            t = features.cat if features.cat is not None else features.cont
            return int(t.shape[0])
        return int(torch.as_tensor(next(iter(features.values()))).shape[0])

    def _lookup(self, features, want_logx=False, zero_buf=None, **kw):
        """ids + fused lookup.  One launch (ctr_embed_fwd_raw) for the unsharded table; the sharded
        table needs the ids on their own (they are bucketed by owner first).  ``zero_buf``: a
        buffer to clear on the way (the tower's per-step accumulators).
        -> (logx or None, lookup outputs...); sets ``self.rows``."""
        if self.world == 1 and hasattr(self.emb, "lookup_features"):
            out = self.emb.lookup_features(self.ids, features, want_logx=want_logx,
                                           zero_buf=zero_buf, **kw)
            self.rows = out[0]
            return out[1:]
        if zero_buf is not None:
            zero_buf.zero_()
        r = self.ids(features, want_logx=want_logx)
        self.rows, logx = r if want_logx else (r, None)
        if getattr(self.emb, "p2p", False):
            self.emb.n_dense = self.dense.numel
            kw["training"] = self._training
        return (logx,) + tuple(self.emb.lookup(self.rows, **kw))

    def _fused_head(self, zs, labels, training, shape):
        """ctr_loss_head over the columns ``zs`` (column 0 = pre-bias first-order sum)."""
        B = zs[0].shape[0]
        if labels is None:
            labels = torch.zeros(B, dtype=torch.float32, device=self.device)
        loss, logits, prob = ops.loss_head(self._head_anchor, self.dense, zs, labels, relu0=True,
                                           grad_scale=1.0 / (B * self.world), training=training)
        return logits.view(shape), prob.view(shape), loss

    def _tower_head(self, tower, X, zs, labels, training, shape, X_lo=None):
        """tower(X) as the last head column: one ctr_tower_mid launch when the tower's shape
        allows, else the per-layer kernels + ctr_loss_head."""
        if not tower.use_mid:
            return self._fused_head(list(zs) + [tower(X, training)], labels, training, shape)
        B = X.shape[0]
        if labels is None:
            labels = torch.zeros(B, dtype=torch.float32, device=self.device)
        loss, logits, prob = ops.tower_head(tower, X, zs, labels, relu0=True,
                                            grad_scale=1.0 / (B * self.world), training=training,
                                            X_lo=X_lo)
        return logits.view(shape), prob.view(shape), loss

    def backward(self, loss):
        # data parallel: the global loss is the mean over all replicas' batches (the fused loss
        # head already folds 1/world into the gradients it emits)
        fused_head = self.fused and hasattr(self, "_uses_fused_head")
        super().backward(loss / self.world if (self.world > 1 and not fused_head) else loss)

    def _sync_dense_grads(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.dense.grad)      # replicated dense weights: sum of per-rank grads

    def _join_tower(self):
        tw = getattr(self, "tower", None)
        if tw is not None:
            tw.join()

    def _dense_step(self, lr_t):
        if getattr(self.emb, "p2p", False):       # all-reduce + Adam over peer memory, no NCCL
            self.emb.dense_step(self.dense, lr_t, self.adam, getattr(self, "tower", None))
            self.dense.lo_fresh.clear()            # the weights moved without their lo halves
        else:
            self._sync_dense_grads()
            self.dense.adam_step(lr_t, self.adam)

    def _apply_gradients(self, lr_t):
        """One GPU, lazy rows: the step's two closing optimiser kernels run side by side - the touched
        rows on the main stream, the dense weights on an optimiser stream that waits for everything
        the main stream has produced so far (every dense gradient) and for the tower's side-stream
        weight-gradient kernels.  The last of the two kernels to finish advances the device schedule
        (``advance_parties``); for a tower with pre-split first-layer operands the dense launch also
        leaves the lo half of those weights for the next step's lookup kernel."""
        emb = self.emb
        if (self.world == 1 and isinstance(emb, ops.FieldEmbedding) and emb.adam_mode == "lazy"
                and self.adam.state is not None and not getattr(emb, "_fused_done", False)
                and getattr(emb, "_fused", None) is None
                and getattr(self, "tower", None) is not None     # FM's 4 dense weights: a fork + join
                                                                 # costs more than the 4 us it hides
                and os.environ.get("CTR_DENSE_ON_SIDE", "1") != "0"):
            main = torch.cuda.current_stream()
            if getattr(self, "_opt_stream", None) is None:
                self._opt_stream = torch.cuda.Stream(device=self.device)
            side = self._opt_stream
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            tw = getattr(self, "tower", None)
            if tw is not None and tw._pending is not None:
                side.wait_event(tw._pending)
                tw._pending = None
            emb.adam_step(self.rows, lr_t, self.adam, parties=2)
            with torch.cuda.stream(side):
                lo = (tw.prefix + ".0.w", tw.w0_lo) if (tw is not None and tw.use_presplit) else None
                self.dense.adam_step(lr_t, self.adam, parties=2, lo=lo)
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
            return
        self.emb.adam_step(self.rows, lr_t, self.adam)     # overlaps the side-stream dW kernels
        if not getattr(self.emb, "p2p", False):
            self._join_tower()                             # (the peer path joins on its side stream)
        self._dense_step(lr_t)


# ============================================================================ FM
class FMModel(_CriteoBase):
    """fm/fm.py:115-170."""
    name = "fm"

    def __init__(self, params):
        super().__init__(params)
        self.dense = ops.DenseParams({"b1": (1,), "head.w": (2, 1), "head.b": (1,)}, self.device)
        self._init_dense(int(params.get("seed", 0)) + 1)

    _uses_fused_head = True

    def logits(self, features, training):
        P = self.dense
        self.rows = self.ids(features)
        E, y1s, y2, _ = self.emb.lookup(self.rows, want_fm=True, want_y1=True)
        y1 = torch.relu(y1s + P["b1"])                                   # fm/fm.py:121
        z = torch.stack([y1, y2], 1)                                      # :131
        return torch.addmm(P["head.b"], z, P["head.w"])                   # :132  [B,1]

    def forward(self, features, labels, training):
        if not self.fused:
            return super().forward(features, labels, training)
        _, E, y1s, y2, _ = self._lookup(features, want_fm=True, want_y1=True)
        return self._fused_head([y1s, y2], labels, training, (-1, 1))     # fm/fm.py:121-149


# ======================================================================== DeepFM
class DeepFMModel(_CriteoBase):
    """deepfm/deepfm.py:73-150 (field-count agnostic; fed the Criteo columns)."""
    name = "deepfm"

    def __init__(self, params):
        super().__init__(params)
        self.layers = list(map(int, str(params["deep_layers"]).split(",")))
        shapes = {"b1": (1,)}
        self._tower_shapes(shapes, "dnn", [self.F * self.D] + self.layers, bn=True)
        shapes.update({"dnn.out.w": (self.layers[-1], 1), "dnn.out.b": (1,), "head.w": (3, 1),
                       "head.b": (1,)})
        frozen = [n for n in shapes if n.endswith((".bn.mean", ".bn.var"))]
        self.dense = ops.DenseParams(shapes, self.device, frozen=frozen)
        self._init_dense(int(params.get("seed", 0)) + 1)
        self.tower = ops.FusedTower(self.dense, "dnn", [self.F * self.D] + self.layers, True,
                                    self.dropout, self.adam, seed=int(params.get("seed", 0)))

    _uses_fused_head = True

    def forward(self, features, labels, training):
        if not self.fused:
            return super().forward(features, labels, training)
        lo = self.tower.use_presplit and (self.world == 1 or getattr(self.emb, "p2p", False) or
                                          getattr(getattr(self.emb, "ops", None), "packed", False))
        # one GPU, unsharded table: ids + lookup + FM terms + the first tower layer in ONE launch
        fuse = bool(lo) and self.world == 1 and isinstance(self.emb, ops.FieldEmbedding) and \
            self.tower.can_fuse_l0(self.F, self.D) and self._batch_size(features) >= 256
        ws = self.tower.begin_step(self._batch_size(features), expect_lo=lo, fused_l0=fuse)
        _, E, y1s, y2, _ = self._lookup(features, zero_buf=ws, want_fm=True, want_y1=True,
                                        **({"want_lo": True} if lo else {}),
                                        **({"tower0": self.tower} if fuse else {}))
        return self._tower_head(self.tower, E, [y1s, y2], labels, training, (-1,),
                                X_lo=self.emb.last_E_lo if lo else None)      # :91,100-129

    def logits(self, features, training):
        P = self.dense
        self.rows = self.ids(features)
        E, y1s, y2, _ = self.emb.lookup(self.rows, want_fm=True, want_y1=True)
        y1 = torch.relu(y1s + P["b1"])                                    # deepfm.py:91
        h = self._tower(E, "dnn", len(self.layers), training)             # :100-107
        y3 = torch.relu(torch.addmm(P["dnn.out.b"], h, P["dnn.out.w"]))   # :108
        z = torch.cat([y1[:, None], y2[:, None], y3], 1)                  # :110
        return torch.addmm(P["head.b"], z, P["head.w"]).reshape(-1)       # :111-112  [B]


# =========================================================================== DCN
class DCNModel(_CriteoBase):
    """dcn/dcn.py:117-190.  linear_feature_columns are built but unused (:122,129-130)."""
    name = "dcn"
    want_w1 = False

    def __init__(self, params):
        super().__init__(params)
        self.layers = list(map(int, str(params["deep_layers"]).split(",")))
        self.L = int(params.get("cross_layers", 4))       # FLAGS.cross_layers, dcn/dcn.py:24,134
        W = self.F * self.D
        shapes = {"cross.w": (self.L, W), "cross.b": (self.L, W)}
        self._tower_shapes(shapes, "dnn", [W] + self.layers, bn=True)
        shapes.update({"head.w": (self.layers[-1] + W, 1), "head.b": (1,)})
        frozen = [n for n in shapes if n.endswith((".bn.mean", ".bn.var"))]
        self.dense = ops.DenseParams(shapes, self.device, frozen=frozen)
        self._init_dense(int(params.get("seed", 0)) + 1)
        # dense(relu) -> BN -> dropout per layer, no final dense layer (dcn/dcn.py:144-149)
        self.tower = ops.FusedTower(self.dense, "dnn", [W] + self.layers, False, self.dropout,
                                    self.adam, seed=int(params.get("seed", 0)))

    _uses_fused_head = True

    def forward(self, features, labels, training):
        if not self.fused:
            return super().forward(features, labels, training)
        P = self.dense
        _, E, _, _, xl = self._lookup(features, want_fm=False, want_y1=False,
                                      cross_w=P["cross.w"], cross_b=P["cross.b"])   # dcn.py:123-142
        h = self.tower(E, training)                                                  # :144-149
        B = E.shape[0]
        if labels is None:
            labels = torch.zeros(B, dtype=torch.float32, device=self.device)
        loss, logits, prob = ops.dcn_head(self._head_anchor, P, h, xl, labels,
                                          grad_scale=1.0 / (B * self.world),
                                          training=training)                        # :151-153,166
        return logits.view(-1, 1), prob.view(-1, 1), loss

    def _init_weight(self, name, shape, g):
        if name == "cross.w":
            return torch.stack([_glorot_normal((shape[1],), g) for _ in range(shape[0])])
        return _glorot_uniform(shape, g)

    def _init_dense(self, seed):
        super()._init_dense(seed)
        g = torch.Generator().manual_seed(seed + 7)
        with torch.no_grad():   # cross bias is glorot-normal too (dcn/dcn.py:140)
            self.dense["cross.b"].copy_(torch.stack(
                [_glorot_normal((self.F * self.D,), g) for _ in range(self.L)]))

    def load_state(self, state):
        st = dict(state)
        if "cross.0.w" in st:
            st["cross.w"] = torch.stack([st[f"cross.{l}.w"] for l in range(self.L)])
            st["cross.b"] = torch.stack([st[f"cross.{l}.b"] for l in range(self.L)])
        super().load_state(st)

    def logits(self, features, training):
        P = self.dense
        _, E, _, _, xl = self._lookup(features, want_fm=False, want_y1=False,
                                      cross_w=P["cross.w"], cross_b=P["cross.b"])
        h = self._tower(E, "dnn", len(self.layers), training)             # dcn.py:144-149
        z = torch.cat([h, xl], 1)                                         # :151
        return torch.addmm(P["head.b"], z, P["head.w"])                   # :152  [B,1]


# ======================================================================= xDeepFM
class XDeepFMModel(_CriteoBase):
    """xdeepfm/xdeepfm.py:123-233: linear (13 log-numerics + 26 one-hots, :82,91,131)
    + CIN (:135-182) + DNN (:184-192).  The DNN branch has its own embedding tables
    because the reference calls input_layer a second time (:185) [TF-sem]; pass
    params['share_embeddings']=True to use one set."""
    name = "xdeepfm"
    needs_logx = True       # the id stage also yields the log-normalised numerics (:82)

    def __init__(self, params):
        super().__init__(params)
        self.layers = list(map(int, str(params["deep_layers"]).split(",")))
        self.cin_layers = list(map(int, str(params["cross_layers"]).split(",")))
        self.cin_precision = params.get("cin_precision", "tf32x3")
        self.share = bool(params.get("share_embeddings", False))
        if self.world > 1 and not self.share:
            raise ValueError("xdeepfm with shard_embedding needs share_embeddings=True: the DNN "
                             "branch's second set of tables (xdeepfm/xdeepfm.py:185) is not sharded")
        seed = int(params.get("seed", 0))
        self.emb_dnn = None if self.share else ops.FieldEmbedding(
            self.lay, self.device, with_w1=False, adam_mode=params.get("embedding_adam", "lazy"),
            seed=seed + 100)
        shapes = {"b1": (1,), "wnum": (len(self.ids.cont_keys),)}
        hp = self.F
        for k, h in enumerate(self.cin_layers):
            shapes[f"cin.{k}.w"] = (self.F * hp, h)
            shapes[f"cin.{k}.b"] = (h,)
            hp = h
        shapes.update({"cin.out.w": (sum(self.cin_layers), 1), "cin.out.b": (1,)})
        self._tower_shapes(shapes, "dnn", [self.F * self.D] + self.layers, bn=True)
        shapes.update({"dnn.out.w": (self.layers[-1], 1), "dnn.out.b": (1,), "head.w": (3, 1),
                       "head.b": (1,)})
        frozen = [n for n in shapes if n.endswith((".bn.mean", ".bn.var"))]
        self.dense = ops.DenseParams(shapes, self.device, frozen=frozen)
        self._init_dense(seed + 1)
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed + 2)
            self.dense["wnum"].copy_(_glorot_uniform((len(self.ids.cont_keys), 1), g).reshape(-1))
        self.tower = ops.FusedTower(self.dense, "dnn", [self.F * self.D] + self.layers, True,
                                    self.dropout, self.adam, seed=seed)

    _uses_fused_head = True

    def forward(self, features, labels, training):
        if not self.fused:
            return super().forward(features, labels, training)
        P = self.dense
        lo = self.tower.use_presplit and self.world == 1 and self.emb_dnn is not None
        ws = self.tower.begin_step(self._batch_size(features), expect_lo=lo)
        want_num = len(self.numeric_linear) > 0
        logx, E, y1s, _, _ = self._lookup(features, want_logx=want_num, zero_buf=ws, want_fm=False,
                                          want_y1=True)
        lin = y1s + logx @ P["wnum"] if want_num else y1s                        # :82
        Ws = [P[f"cin.{k}.w"] for k in range(len(self.cin_layers))]
        bs = [P[f"cin.{k}.b"] for k in range(len(self.cin_layers))]
        pooled = ops.cin(E, self.F, self.D, Ws, bs, self.cin_precision)          # :135-181
        cin_y = torch.relu(torch.addmm(P["cin.out.b"], pooled, P["cin.out.w"])).view(-1)  # :182
        if self.emb_dnn is not None:
            Ed = self.emb_dnn.lookup(self.rows, want_fm=False, want_y1=False,
                                     **({"want_lo": True} if lo else {}))[0]       # :185
        else:
            Ed = E
        return self._tower_head(self.tower, Ed, [lin, cin_y], labels, training, (-1, 1),
                                X_lo=self.emb_dnn.last_E_lo if lo else None)      # :131,188-212

    def load_state(self, state):
        super().load_state(state)
        if self.emb_dnn is not None and "emb_dnn" in state:
            self.emb_dnn.load(state["emb_dnn"])

    def logits(self, features, training):
        P = self.dense
        want_num = len(self.numeric_linear) > 0
        if want_num:
            self.rows, logx = self.ids(features, want_logx=True)
        else:
            self.rows, logx = self.ids(features), None
        E, y1s, _, _ = self.emb.lookup(self.rows, want_fm=False, want_y1=True)
        lin = y1s + P["b1"]
        if want_num:
            lin = lin + logx @ P["wnum"]                                        # :82
        linear_y = torch.relu(lin)                                              # :131
        Ws = [P[f"cin.{k}.w"] for k in range(len(self.cin_layers))]
        bs = [P[f"cin.{k}.b"] for k in range(len(self.cin_layers))]
        pooled = ops.cin(E, self.F, self.D, Ws, bs, self.cin_precision)         # :135-181
        cin_y = torch.relu(torch.addmm(P["cin.out.b"], pooled, P["cin.out.w"]))  # :182
        if self.emb_dnn is not None:
            Ed = self.emb_dnn.lookup(self.rows, want_fm=False, want_y1=False)[0]  # :185
        else:
            Ed = E
        h = self._tower(Ed, "dnn", len(self.layers), training)                  # :188-191
        dnn_y = torch.relu(torch.addmm(P["dnn.out.b"], h, P["dnn.out.w"]))      # :192
        z = torch.cat([linear_y[:, None], cin_y, dnn_y], 1)                     # :194
        return torch.addmm(P["head.b"], z, P["head.w"])                         # :195  [B,1]

    def _apply_gradients(self, lr_t):
        self.emb.adam_step(self.rows, lr_t, self.adam)
        if self.emb_dnn is not None:
            self.emb_dnn.adam_step(self.rows, lr_t, self.adam)
        self._join_tower()
        self._dense_step(lr_t)


# =========================================================================== DIN
DIN_ITEMS, DIN_CATES = 63002, 802          # din/din.py:88-90 (hard-coded there)
DIN_ATT_LAYERS = [80, 40]                  # din/din.py:85
DIN_MLP_LAYERS = [100, 50, 20]             # din/din.py:86


class DINModel(_ModelBase):
    """din/din.py:83-180.  The two tables (i_id, i_cate) live in one FieldEmbedding
    (sub-table 0 and 1, same width); i_item is its first-order vector."""
    name = "din"

    def __init__(self, params):
        super().__init__(params)
        E = int(params["embedding_size"])
        self.E = E
        n_items = int(params.get("din_items", DIN_ITEMS))
        n_cates = int(params.get("din_cates", DIN_CATES))
        cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("i_id", n_items), E),
                fc.embedding_column(fc.categorical_column_with_hash_bucket("i_cate", n_cates), E)]
        self.lay = fc.Layout(cols, ["i_id", "i_cate"], [n_items, n_cates],
                             [0, n_items, n_items + n_cates], E)
        seed = int(params.get("seed", 0))
        # planar tables: the activation-unit kernels (din.cu) index rows with stride E
        self.emb = ops.FieldEmbedding(self.lay, self.device, with_w1=True, w1_fields=0b01,
                                      adam_mode=params.get("embedding_adam", "exact_tf"), seed=seed,
                                      record=False)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():   # glorot_normal tables, zero item bias (din/din.py:88-90)
            self.emb.table[:n_items].copy_(_glorot_normal((n_items, E), g))
            self.emb.table[n_items:].copy_(_glorot_normal((n_cates, E), g))
            self.emb.w1.zero_()
        shapes = {}
        for name in ("att_iid", "att_cat"):
            self._tower_shapes(shapes, name, [4 * E] + DIN_ATT_LAYERS + [1], bn=False)
        self._tower_shapes(shapes, "mlp", [3 * E] + DIN_MLP_LAYERS, bn=False)
        shapes.update({"mlp.out.w": (DIN_MLP_LAYERS[-1], 1), "mlp.out.b": (1,)})
        self.dense = ops.DenseParams(shapes, self.device)
        self._init_dense(seed + 1)
        self.n_items, self.n_cates = n_items, n_cates
        self.rows = None
        self.seed = seed
        self.fused = bool(params.get("fused_tower", True))
        # target ids -> rows of the concatenated table, range-checked on the device
        self.ids = ops.IdPipeline(self.lay, self.device)
        # MLP 100-50-20 (ReLU + dropout, no BN) + dense(1) without activation (din/din.py:133-139)
        self.tower = ops.FusedTower(self.dense, "mlp", [3 * E] + DIN_MLP_LAYERS, True, self.dropout,
                                    self.adam, seed=seed, bn=False, out_relu=False, layer_base=8)
        # logit = mlp_out + i_b (:140) through ctr_loss_head: constant head weights [1, 1], bias 0;
        # their "gradients" land in a scratch buffer that nothing reads
        self._head_w = torch.ones(2, dtype=torch.float32, device=self.device)
        self._head_b = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._head_scratch = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._head_anchor = torch.zeros((), device=self.device, requires_grad=True)

    def load_state(self, state):
        super().load_state(state)
        if "i_id" in state:
            tab = torch.cat([state["i_id"], state["i_cate"]], 0)
            w1 = torch.cat([state["i_item"].reshape(-1),
                            torch.zeros(self.lay.rows[1], dtype=state["i_item"].dtype)])
            self.emb.load(tab, w1)

    def att_opts(self, unit, training):
        """ctr_din_opts of attention unit ``unit``: the dropout of din/din.py:118 (training only)
        and the id range check against the unit's table."""
        return ops.din_opts(unit, self.lay.rows[unit], self.dropout if training else 0.0, self.seed,
                            self.adam.state_ptr, self.ids.status)

    def _att(self, prefix, field, hist, query, training):
        P = self.dense
        return ops.din_attention(self.emb, field, hist, query, P[f"{prefix}.0.w"], P[f"{prefix}.0.b"],
                                 P[f"{prefix}.1.w"], P[f"{prefix}.1.b"], P[f"{prefix}.2.w"],
                                 P[f"{prefix}.2.b"], self.att_opts(field, training))

    def _net(self, features, training):
        """-> (net [B, 3E], i_b [B]): target embeddings + the two attended histories (:91-131)."""
        dev = self.device
        h_iid = torch.as_tensor(features["u_iid_seq"]).to(dev, non_blocking=True).to(torch.int32)
        h_cat = torch.as_tensor(features["u_icat_seq"]).to(dev, non_blocking=True).to(torch.int32)
        self.rows = self.ids({"i_id": torch.as_tensor(features["i_id"]).reshape(-1, 1),
                              "i_cate": torch.as_tensor(features["i_cate"]).reshape(-1, 1)})
        self.hist = (h_iid, h_cat)
        Ecat, i_b, _, _ = self.emb.lookup(self.rows, want_fm=False, want_y1=True)   # :91-99
        E = self.E
        pkg_emb, pkgc_emb = Ecat[:, :E], Ecat[:, E:]
        pkg_h = self._att("att_iid", 0, h_iid, pkg_emb, training)                   # :127
        pkgc_h = self._att("att_cat", 1, h_cat, pkgc_emb, training)                 # :128
        return torch.cat([pkg_emb, pkg_h, pkgc_h], 1), i_b                          # :131

    def forward(self, features, labels, training):
        if not self.fused:
            return super().forward(features, labels, training)
        net, i_b = self._net(features, training)
        y = self.tower(net, training)                                               # :133-139
        if labels is None:
            labels = torch.zeros(net.shape[0], dtype=torch.float32, device=self.device)
        loss, logits, prob = ops.loss_head(
            self._head_anchor, self.dense, [y, i_b], labels,
            hw=(self._head_w, self._head_scratch[0:2]), hb=(self._head_b, self._head_scratch[2:3]),
            b1=(None, None), relu0=False, training=training)                        # :140-141,166
        return logits, prob, loss

    def logits(self, features, training):
        """The torch (cuBLAS) MLP of params['fused_tower'] = False; no dropout there."""
        if training and self.dropout > 0.0:
            raise NotImplementedError("the torch MLP path has no device-side dropout stream; use "
                                      "the default fused_tower=True")
        P = self.dense
        net, i_b = self._net(features, training)
        net = self._tower(net, "mlp", len(DIN_MLP_LAYERS), training, bn=False)      # :133-137
        out = torch.addmm(P["mlp.out.b"], net, P["mlp.out.w"])                      # :139
        return out.reshape(-1) + i_b                                                # :140

    def _apply_gradients(self, lr_t):
        if self.emb.adam_mode == "exact_tf":
            self.emb.adam_step(self.rows, lr_t, self.adam)
        else:
            h_iid, h_cat = self.hist
            touched = torch.cat([self.rows.reshape(-1), h_iid.reshape(-1),
                                 (h_cat + self.n_items).reshape(-1)])
            self.emb.adam_step(touched, lr_t, self.adam)
        self.tower.join()
        self.dense.adam_step(lr_t, self.adam)

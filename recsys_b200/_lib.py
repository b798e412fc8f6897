"""ctypes binding of libctr_b200.so (include/ctr_b200.h).

The shared library is the product: there is no Python or CPU fallback.  If it is
missing (not built) the import of any compute symbol raises; if the current CUDA
device is not sm_100 every compute call returns CTR_ERR_ARCH and ``check``
raises ``RuntimeError`` with ``ctr_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libctr_b200.so")

_lib = None

c_f = C.c_void_p      # device pointers are passed as integers (tensor.data_ptr())
c_i = C.c_int
c_i64 = C.c_int64
c_u64 = C.c_uint64
c_fl = C.c_float


class BnDrop(C.Structure):
    """ctr_bn_drop (include/ctr_b200.h)."""
    _fields_ = [("sums", C.c_void_p), ("mean", C.c_void_p), ("var", C.c_void_p),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("state", C.c_void_p),
                ("eps", C.c_float), ("p_drop", C.c_float), ("seed", C.c_uint32),
                ("layer", C.c_uint32), ("enabled", C.c_int32), ("pad_", C.c_int32)]


class GradSrc(C.Structure):
    """ctr_grad_src (include/ctr_b200.h)."""
    _fields_ = [("G", C.c_void_p), ("a", C.c_void_p), ("sums", C.c_void_p), ("mean", C.c_void_p),
                ("var", C.c_void_p), ("gamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("dgamma", C.c_void_p), ("ldg", C.c_int32), ("lda", C.c_int32), ("eps", C.c_float),
                ("kind", C.c_int32), ("train", C.c_int32), ("pad_", C.c_int32)]


_ML = 4   # CTR_TOWER_MID_MAX_LAYERS


class TowerMidArgs(C.Structure):
    """ctr_tower_mid_args (include/ctr_b200.h)."""
    _P4 = C.c_void_p * _ML
    _fields_ = [("L", C.c_int32), ("C", C.c_int32), ("relu0", C.c_int32), ("training", C.c_int32),
                ("H", C.c_int32 * _ML),
                ("W", _P4), ("b", _P4), ("gamma", _P4), ("beta", _P4), ("mean", _P4), ("var", _P4),
                ("act", _P4), ("stats", _P4),
                ("w_out", C.c_void_p), ("b_out", C.c_void_p), ("state", C.c_void_p),
                ("eps", C.c_float), ("p_drop", C.c_float), ("seed", C.c_uint32),
                ("grad_scale", C.c_float),
                ("z", C.c_void_p * 3), ("hw", C.c_void_p), ("hb", C.c_void_p), ("b1", C.c_void_p),
                ("labels", C.c_void_p), ("y_out", C.c_void_p), ("logits", C.c_void_p),
                ("prob", C.c_void_p), ("loss", C.c_void_p),
                ("dz", C.c_void_p * 3), ("dhw", C.c_void_p), ("dhb", C.c_void_p),
                ("db1", C.c_void_p), ("dw_out", C.c_void_p), ("db_out", C.c_void_p),
                ("dgamma", _P4), ("dbeta", _P4), ("dbias", _P4), ("dn", _P4), ("dpre", _P4),
                ("dpre0_lo", C.c_void_p), ("pre0", C.c_void_p), ("barrier", C.c_void_p), ("timing", C.c_void_p),
                ("stats0_part", C.c_void_p), ("n_stats0_part", C.c_int32), ("pad_", C.c_int32)]


class DinOpts(C.Structure):
    """ctr_din_opts (include/ctr_b200.h)."""
    _fields_ = [("state", C.c_void_p), ("p_drop", C.c_float), ("seed", C.c_uint32),
                ("unit", C.c_uint32), ("table_rows", C.c_int32), ("status", C.c_void_p)]


class P2PCtx(C.Structure):
    """ctr_p2p_ctx (include/ctr_b200.h)."""
    _fields_ = [("peer", C.c_void_p * 8), ("me", C.c_int32), ("G", C.c_int32),
                ("capacity", C.c_int32), ("record_floats", C.c_int32),
                ("off_req_flag", C.c_int64), ("off_req_cnt", C.c_int64), ("off_resp_flag", C.c_int64),
                ("off_grad_flag", C.c_int64), ("off_dense_flag", C.c_int64),
                ("off_req_ids", C.c_int64), ("off_resp", C.c_int64), ("off_grad", C.c_int64),
                ("off_dense", C.c_int64), ("off_counts", C.c_int64), ("off_done", C.c_int64),
                ("off_inv", C.c_int64), ("off_sent", C.c_int64), ("n_dense", C.c_int64), ("spin_limit_ms", C.c_int32), ("pad_", C.c_int32)]


class FieldDesc(C.Structure):
    """ctr_field_desc (include/ctr_b200.h)."""
    _fields_ = [("kind", C.c_int32), ("src", C.c_int32), ("n_rows", C.c_int32),
                ("row_offset", C.c_int32), ("bnd_begin", C.c_int32), ("bnd_count", C.c_int32),
                ("log_offset", C.c_float), ("pad_", C.c_int32)]


# name -> (restype, argtypes); every symbol include/ctr_b200.h declares.
SIGNATURES = {
    "ctr_version": (c_i, []),
    "ctr_last_error": (C.c_char_p, []),
    "ctr_device_check": (c_i, []),
    "ctr_set_option": (c_i, [C.c_char_p, c_i]),
    "ctr_criteo_rows": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_f]),
    "ctr_criteo_rows_bg": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_i, c_f]),
    "ctr_hash_strings": (c_i, [c_f, c_f, c_i64, c_f, c_f, c_f, c_f, c_f]),
    "ctr_hash_int64": (c_i, [c_f, c_i64, C.c_int32, c_f, c_f]),
    "ctr_hash_slots": (c_i, [c_f, c_i, c_f, c_i64, c_i, c_f, c_f, c_f]),
    "ctr_tfrecord_scan": (c_i64, [c_f, c_i64, c_i, c_f, c_f, c_i64]),
    "ctr_masked_crc32c": (C.c_uint32, [c_f, c_i64]),
    "ctr_criteo_parse": (c_i, [c_f, c_f, c_f, c_i64, c_i, c_f, c_f, c_f, c_f, c_i]),
    "ctr_din_parse": (c_i64, [c_f, c_f, c_f, c_i64, c_i, c_i64, c_f, c_f, c_f, c_f, c_f]),
    "ctr_embed_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_u64, c_f, c_f, c_f, c_f, c_f, c_f, c_i,
                            c_f, c_f, c_i64, c_i64, c_f]),
    "ctr_embed_fwd_raw": (c_i, [c_f, c_f, c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_f, c_f, c_f, c_i, c_i, c_i,
                                c_u64, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_i64, c_i64, c_f, c_i64,
                                c_f]),
    "ctr_embed_tower_fwd": (c_i, [c_f, c_f, c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_u64,
                                  c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_f, c_f, c_f, c_i, c_f, c_f,
                                  c_f, c_i64, c_f]),
    "ctr_embed_tower_timing": (c_i, [c_f]),
    "ctr_tower_embed_bwd": (c_i, [c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_f, c_f, c_f, c_u64,
                                  C.POINTER(C.c_int64), c_i, c_i, c_i, c_f, c_f, c_i64, c_i64, c_f]),
    "ctr_embed_bwd": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_u64, C.POINTER(C.c_int64), c_i,
                            c_i, c_i, c_f, c_f, c_i64, c_i64, c_f]),
    "ctr_count_rows": (c_i, [c_f, c_i64, c_i, c_f, c_i64, c_f]),
    "ctr_embed_bwd_adam": (c_i, [c_f, c_f, c_f, c_f, c_f, c_u64, C.POINTER(C.c_int64), c_i, c_i, c_i,
                                 c_f, c_i64, c_fl, c_fl, c_fl, c_fl, c_f, c_f]),
    "ctr_dcn_cross_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f]),
    "ctr_dcn_cross_bwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f]),
    "ctr_adam_tick": (c_i, [c_f, c_fl, c_fl, c_fl, c_f]),
    "ctr_adam_dense": (c_i, [c_f, c_f, c_f, c_f, c_i64, c_fl, c_fl, c_fl, c_fl, c_i, c_f, c_i, c_f]),
    "ctr_adam_rows": (c_i, [c_f, c_i64, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                            C.c_int32, c_fl, c_fl, c_fl, c_fl, c_f, c_i64, c_i64, c_i64, c_f]),
    "ctr_adam_rows_bf": (c_i, [c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                               C.c_int32, c_fl, c_fl, c_fl, c_fl, c_f, c_i64, c_i64, c_i64, c_i, c_f]),
    "ctr_adam_rows_ex": (c_i, [c_f, c_i64, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                               C.c_int32, c_fl, c_fl, c_fl, c_fl, c_f, c_i64, c_i64, c_i64, c_i, c_f]),
    "ctr_adam_dense_ex": (c_i, [c_f, c_f, c_f, c_f, c_i64, c_fl, c_fl, c_fl, c_fl, c_i, c_f, c_i,
                                c_f, c_i64, c_i64, c_f]),
    "ctr_din_att_fwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_i, c_f, c_f, c_i, c_f, c_f,
                              c_f, c_f, C.POINTER(DinOpts), c_f]),
    "ctr_din_dropout_mask": (c_i, [C.POINTER(DinOpts), c_i, c_i64, c_i, c_f, c_f]),
    "ctr_din_workspace_bytes": (c_i64, [c_i, c_i, c_i]),
    "ctr_din_att_bwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_i, c_f, c_f, c_i, c_f, c_f,
                              c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i64,
                              C.POINTER(DinOpts), c_f]),
    "ctr_cin_workspace_bytes": (c_i64, [c_i, c_i, c_i, c_i, c_i, c_i]),
    "ctr_cin_layer_fwd": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_i,
                                c_f, c_i64, c_f]),
    "ctr_cin_layer_bwd": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f,
                                c_f, c_f, c_i, c_f, c_i64, c_f]),
    "ctr_tower_layer_fwd": (c_i, [c_f, c_i, c_i, C.POINTER(BnDrop), c_f, c_f, c_i, c_f, c_i, c_f, c_i,
                                  c_i, c_f]),
    "ctr_bn_drop_apply": (c_i, [c_f, c_i, C.POINTER(BnDrop), c_f, c_i, c_f]),
    "ctr_bn_drop_apply_bwd": (c_i, [c_f, c_i, c_f, c_i, C.POINTER(BnDrop), c_f, c_f, c_f, c_i, c_f]),
    "ctr_dcn_head": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_f, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                           c_fl, c_f]),
    "ctr_tower_layer_bwd_data": (c_i, [C.POINTER(GradSrc), c_i, c_f, c_i, C.POINTER(BnDrop), c_f,
                                       c_f, c_i, c_f, c_f, c_i, c_f]),
    "ctr_tower_dpre": (c_i, [C.POINTER(GradSrc), c_i, c_f, c_i, c_f, c_i, c_f]),
    "ctr_tower_layer_bwd_weights": (c_i, [c_f, c_i, c_i, C.POINTER(BnDrop), C.POINTER(GradSrc), c_i,
                                          c_f, c_f, c_i, c_f]),
    "ctr_loss_head": (c_i, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_i, c_i, c_f, c_f, c_f,
                            c_f, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_fl, c_f]),
    "ctr_split_lo": (c_i, [c_f, c_f, c_i64, c_f]),
    "ctr_tower_gemm_presplit": (c_i, [c_i, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_i, c_f]),
    "ctr_tower_mid": (c_i, [C.POINTER(TowerMidArgs), c_i, c_f]),
    "ctr_shard_bucket": (c_i, [c_f, c_i64, c_i, c_i, c_f, c_f, c_f, c_f]),
    "ctr_gather_rows": (c_i, [c_f, c_f, c_f, c_i64, c_i, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_f]),
    "ctr_scatter_add_rows": (c_i, [c_f, c_f, c_f, c_i64, c_i, c_f, c_f, c_i64, c_i64, c_i64, c_i64,
                                   c_f]),
    "ctr_p2p_alloc": (c_i, [c_i64, C.POINTER(C.c_void_p), c_f]),
    "ctr_p2p_open": (c_i, [c_f, C.POINTER(C.c_void_p)]),
    "ctr_p2p_close": (c_i, [c_f]),
    "ctr_p2p_free": (c_i, [c_f]),
    "ctr_p2p_bucket_send": (c_i, [c_f, c_i64, C.POINTER(P2PCtx), c_f, c_f]),
    "ctr_p2p_gather_reply": (c_i, [c_f, c_i64, c_i, c_i, c_i, C.POINTER(P2PCtx), c_f]),
    "ctr_embed_fwd_p2p": (c_i, [c_f, c_i, c_i, c_i, c_u64, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_f,
                                c_f, C.POINTER(P2PCtx), c_f]),
    "ctr_p2p_wait": (c_i, [C.POINTER(P2PCtx), c_i, c_f]),
    "ctr_p2p_grad_send": (c_i, [c_f, c_f, c_f, c_f, c_f, c_u64, c_i, c_i, c_i, C.POINTER(P2PCtx), c_f]),
    "ctr_p2p_scatter_adam": (c_i, [c_f, c_i64, c_i, c_i, c_i, c_fl, c_fl, c_fl, c_fl, c_f,
                                   C.POINTER(P2PCtx), c_f]),
    "ctr_p2p_dense_push": (c_i, [c_f, c_i64, C.POINTER(P2PCtx), c_f]),
    "ctr_p2p_adam_dense": (c_i, [c_f, c_f, c_f, c_f, c_i64, c_fl, c_fl, c_fl, c_fl, c_f, c_i,
                                 C.POINTER(P2PCtx), c_f]),
    "ctr_p2p_status": (c_i, [C.POINTER(P2PCtx), c_f, c_f]),
    "ctr_cin_pool": (c_i, [c_f, c_i, c_i, c_i, c_f, c_i, c_f]),
    "ctr_cin_dpre": (c_i, [c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_f, c_f]),
    "ctr_transpose_fd": (c_i, [c_f, c_i, c_i, c_i, c_f, c_i, c_f]),
    "ctr_transpose_df_add": (c_i, [c_f, c_i, c_i, c_i, c_i, c_f, c_f]),
}


def load():
    """Load libctr_b200.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "recsys_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    # tuning options from the environment: CTR_OPTIONS="name=value,name=value" (ctr_set_option)
    for kv in filter(None, os.environ.get("CTR_OPTIONS", "").split(",")):
        name, _, val = kv.partition("=")
        if lib.ctr_set_option(name.strip().encode(), int(val)) != 0:
            raise RuntimeError("CTR_OPTIONS: " + lib.ctr_last_error().decode())
    return lib


def last_error() -> str:
    return load().ctr_last_error().decode()


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libctr_b200 error %d: %s" % (rc, last_error()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream

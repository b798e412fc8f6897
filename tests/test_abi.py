"""CPU: the C-ABI library builds, loads, and exports every symbol include/ctr_b200.h
declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ctr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = _declared()
    for n in ("ctr_embed_fwd", "ctr_embed_bwd", "ctr_criteo_rows", "ctr_hash_strings",
              "ctr_dcn_cross_fwd", "ctr_dcn_cross_bwd", "ctr_din_att_fwd", "ctr_din_att_bwd",
              "ctr_cin_layer_fwd", "ctr_cin_layer_bwd", "ctr_adam_rows", "ctr_adam_dense"):
        assert n in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), "libctr_b200.so does not export %s" % name


def test_python_binding_covers_the_header(built_lib):
    from recsys_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.ctr_version() >= 100


def test_no_gpu_means_arch_error_not_fallback(built_lib):
    import torch
    from recsys_b200 import _lib
    lib = _lib.load()
    if torch.cuda.is_available():
        assert lib.ctr_device_check() == 0
    else:
        assert lib.ctr_device_check() == -3          # CTR_ERR_ARCH
        assert "CUDA" in _lib.last_error() or "sm_100" in _lib.last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "recsys_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f

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


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of every struct that crosses the C ABI, as gcc lays them out from
    include/ctr_b200.h, against the ctypes mirrors in recsys_b200/_lib.py."""
    import subprocess
    from recsys_b200 import _lib
    checks = {
        "ctr_field_desc": (_lib.FieldDesc, ["kind", "n_rows", "bnd_count", "log_offset"]),
        "ctr_bn_drop": (_lib.BnDrop, ["sums", "gamma", "state", "eps", "seed", "enabled"]),
        "ctr_grad_src": (_lib.GradSrc, ["G", "a", "dgamma", "ldg", "eps", "train"]),
        "ctr_din_opts": (_lib.DinOpts, ["state", "p_drop", "seed", "unit", "table_rows", "status"]),
        "ctr_tower_mid_args": (_lib.TowerMidArgs, ["L", "H", "W", "act", "stats", "w_out", "eps", "seed",
                                                   "grad_scale", "z", "labels", "loss", "dz", "dw_out",
                                                   "dgamma", "dpre", "dpre0_lo", "pre0", "barrier",
                                                   "timing"]),
    }
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "ctr_b200.h"', "int main(void) {"]
    for cname, (_, fields) in checks.items():
        lines.append('printf("%s.sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in fields:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(l.rsplit(" ", 1) for l in out.strip().splitlines())
    for cname, (cls, fields) in checks.items():
        assert int(got[cname + ".sizeof"]) == ctypes.sizeof(cls), cname
        for f in fields:
            assert int(got["%s.%s" % (cname, f)]) == getattr(cls, f).offset, (cname, f)


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/ctr_b200.h against recsys_b200/_lib.SIGNATURES: argument count
    and, per argument, the ctypes class (pointer / int / int64 / uint64 / float) - a mismatch
    here would corrupt the call frame silently."""
    from recsys_b200 import _lib
    src = open(os.path.join(ROOT, "include", "ctr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|uint32_t|const char\s*\*)\s+(ctr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) >= 30

    def kind(arg):
        arg = " ".join(arg.split())
        if arg in ("void", ""):
            return None
        if "*" in arg or arg.startswith("ctr_stream_t"):
            return "ptr"
        t = arg.rsplit(" ", 1)[0].replace("const ", "").strip()
        return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "float": "f32",
                "uint32_t": "u32"}[t]

    def ckind(ct):
        if ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, "contents"):
            return "ptr"
        return {ctypes.c_int: "i32", ctypes.c_int32: "i32", ctypes.c_int64: "i64",
                ctypes.c_uint64: "u64", ctypes.c_float: "f32", ctypes.c_uint32: "u32"}[ct]

    seen = set()
    for name, args in protos:
        kinds = [k for k in (kind(a) for a in args.split(",")) if k is not None]
        res, argtypes = _lib.SIGNATURES[name]
        assert [ckind(t) for t in argtypes] == kinds, (name, kinds, [ckind(t) for t in argtypes])
        seen.add(name)
    assert seen == set(_lib.SIGNATURES)

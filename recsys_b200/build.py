"""In-tree build of libctr_b200.so with nvcc for sm_100a (no JIT cache, no torch
extension machinery: the .so is a plain C-ABI library loaded through ctypes).

Every ``csrc/*.cu`` is its own translation unit (no relocatable device code), so the
objects are compiled in parallel into ``build/obj`` and only the stale ones are rebuilt;
the link step is one ``nvcc -shared``."""
from __future__ import annotations

import glob
import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libctr_b200.so")
OBJ = os.path.join(HERE, "..", "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# debug knobs (getenv overrides of tcgen05 descriptor bits / schedules) are compiled in only
# with CTR_DEBUG_KNOBS=1 in the environment of the build
if os.environ.get("CTR_DEBUG_KNOBS") == "1":
    NVCC_FLAGS.append("-DCTR_DEBUG_KNOBS=1")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


_INC = re.compile(r'^\s*#include\s+"([^"]+)"', re.M)


def _deps(path, seen=None):
    """Transitive quoted includes of ``path`` (files that exist)."""
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in _INC.findall(f.read()):
            _deps(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)
    return seen


def _obj(src):
    tag = "dbg_" if "-DCTR_DEBUG_KNOBS=1" in NVCC_FLAGS else ""
    return os.path.join(OBJ, tag + os.path.basename(src)[:-3] + ".o")


def _stale(src):
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in _deps(src))


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(_stale(s) or os.path.getmtime(_obj(s)) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    todo = [s for s in sources() if force or _stale(s)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(src), src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, cmd, res

    with ThreadPoolExecutor(max_workers=min(len(todo) or 1, os.cpu_count() or 4)) as ex:
        for src, cmd, res in ex.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-6000:]))
            if verbose:
                print(res.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + \
        [_obj(s) for s in sources()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

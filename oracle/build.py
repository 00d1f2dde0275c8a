"""Build/load the C half of the oracle (oracle/chain.c).  TEST INFRASTRUCTURE ONLY.

`oracle/_ref/` is reserved for binaries compiled from the reference's own
sources; the reference is pure Python (no C/C++ at all), so
nothing is ever built there.  Our own restatement builds into `oracle/_build/`.
"""

import ctypes
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "chain.c"
_OUT = _HERE / "_build" / "liboracle_chain.so"
_lib = None


def build_chain_lib(force: bool = False) -> Path:
    if _OUT.exists() and not force and _OUT.stat().st_mtime >= _SRC.stat().st_mtime:
        return _OUT
    _OUT.parent.mkdir(parents=True, exist_ok=True)
    tmp = _OUT.with_suffix(f".{os.getpid()}.tmp")
    # -ffp-contract=off: only the explicit fmaf() calls fuse, so the stated
    # arithmetic is exactly what runs.  -mfma makes __builtin_fmaf one instruction.
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off",
           "-fno-fast-math", str(_SRC), "-o", str(tmp), "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, _OUT)
    return _OUT


def load_chain_lib():
    global _lib
    if _lib is not None:
        return _lib
    path = build_chain_lib()
    lib = ctypes.CDLL(str(path))
    f32p = ctypes.POINTER(ctypes.c_float)
    i64p = ctypes.POINTER(ctypes.c_int64)
    i32p = ctypes.POINTER(ctypes.c_int)
    i64 = ctypes.c_int64
    lib.oracle_dot_chain.restype = ctypes.c_float
    lib.oracle_dot_chain.argtypes = [f32p, f32p, ctypes.c_int]
    lib.oracle_scores_chain.restype = None
    lib.oracle_scores_chain.argtypes = [f32p, i64, ctypes.c_int, f32p, i64, f32p]
    lib.oracle_search_chain.restype = None
    lib.oracle_search_chain.argtypes = [f32p, i64, ctypes.c_int, f32p, i64, ctypes.c_int, i64, f32p, i64p]
    lib.oracle_heap_addn.restype = None
    lib.oracle_heap_addn.argtypes = [ctypes.c_void_p, i32p, i64, ctypes.c_int, f32p, i64, i64]
    lib.oracle_heap_finish.restype = None
    lib.oracle_heap_finish.argtypes = [ctypes.c_void_p, i32p, i64, ctypes.c_int, f32p, i64p]
    lib.oracle_hit_size.restype = ctypes.c_int
    lib.oracle_normalize_l2.restype = None
    lib.oracle_normalize_l2.argtypes = [f32p, i64, ctypes.c_int]
    lib.oracle_merge_topk.restype = None
    lib.oracle_merge_topk.argtypes = [f32p, i64p, ctypes.c_int, i64, ctypes.c_int, f32p, i64p]
    _lib = lib
    return lib

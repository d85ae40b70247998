"""ctypes binding of the C ABI (include/text2loc_b200.h) and the in-tree nvcc build.

The shared library is built in-tree (text2loc_b200/libtext2loc_b200.so) by ``build()`` with
``nvcc -gencode arch=compute_100a,code=sm_100a`` and travels to the GPU box with the repo.
There is no fallback: if the library is missing or does not load, importing the engine
raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libtext2loc_b200.so")
SOURCES = ["api.cu", "linear.cu", "geometry.cu", "pointnet.cu", "sa_obj2.cu", "rowops.cu", "search.cu", "bookkeeping.cu", "synthgen.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]

# every symbol include/text2loc_b200.h declares: (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    "t2l_create": (c_int, [c_int, POINTER(c_void_p)]),
    "t2l_destroy": (None, [_P]),
    "t2l_last_error": (c_char_p, [_P]),
    "t2l_version": (c_int, []),
    "t2l_set_weight": (c_int, [_P, c_char_p, POINTER(c_float), c_int, c_int]),
    "t2l_finalize_weights": (c_int, [_P]),
    "t2l_reserve": (c_int, [_P, c_int, c_int, c_int, c_int, c_int]),
    "t2l_encode_cells": (c_int, [_P, _P, _P, POINTER(c_int32), c_int, _P, _P]),
    "t2l_encode_objects_debug": (c_int, [_P, _P, POINTER(c_int32), c_int] + [_P] * 10 + [_P]),
    "t2l_encode_text": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "t2l_encode_text_tokens": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "t2l_encode_text_tokens_f16": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "t2l_encode_text_sentences": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "t2l_fine_offsets": (c_int, [_P, _P, _P, POINTER(c_int32), c_int, _P, c_int, c_int, _P, _P]),
    "t2l_fine_encode_objects": (c_int, [_P, _P, _P, POINTER(c_int32), c_int, _P, _P]),
    "t2l_fine_encode_hints": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "t2l_fine_match": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "t2l_db_build": (c_int, [_P, _P, c_int64, c_int64, _P]),
    "t2l_search_topk": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P]),
    "t2l_search_topk_accumulate": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P]),
    "t2l_synth_cells": (c_int, [_P, ctypes.c_uint64, c_int64, c_int, c_int, _P, _P, _P]),
    "t2l_search_topk_exact": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "t2l_merge_topk": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "t2l_merge_topk_packed": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "t2l_topk_accuracy": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P, POINTER(c_int32), c_int, POINTER(c_double), c_int, _P, _P, _P, _P]),
    "t2l_launch_count": (c_int64, [_P]),
    "t2l_debug_linear": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "t2l_debug_linear_f16": (c_int, [_P, _P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "t2l_debug_mha": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "t2l_debug_sa_bisect": (c_int, [_P, c_int]),
    "t2l_debug_mha_cross": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "t2l_debug_mha_cells": (c_int, [_P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P]),
    "t2l_debug_linear_f16_residual": (c_int, [_P, _P, c_int, _P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
}


def _nvcc() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isfile(p) or os.sep not in p):
            return p
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "text2loc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into the in-tree shared library (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(obj)
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


_LIB = None


def load() -> ctypes.CDLL:
    """dlopen the engine; raises (never falls back) when it is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(python -c 'import __graft_entry__ as g; g.build()').  text2loc_b200 has no CPU path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib

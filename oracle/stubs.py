"""sys.modules shims so the reference's own Python imports unchanged.  TEST INFRASTRUCTURE.

Missing here: torch_geometric / torch_cluster / torch_scatter, nltk,
easydict, matplotlib, plus numpy-1 module paths removed in numpy 2.  This installs the
minimum surface the hot-path modules touch; the PyG ops are oracle/pyg_ops.py.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import pyg_ops

_INSTALLED = False


class EasyDict(dict):
    """dict with attribute access (enough of easydict.EasyDict for pointnet2.py:94-100)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def sent_tokenize(text: str):
    """nltk.tokenize.sent_tokenize stand-in: split after '.', which is exact for the templated
    hints 'The pose is <dir> of a <colour> <class>.' (dataloading/kitti360pose/base.py:60-68)."""
    parts = [p.strip() for p in text.split(".")]
    return [p + "." for p in parts if p]


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    if "easydict" not in sys.modules:
        _module("easydict", EasyDict=EasyDict)
    tok = _module("nltk.tokenize", sent_tokenize=sent_tokenize)
    _module("nltk", tokenize=tok)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        plt = _module("matplotlib.pyplot")
        _module("matplotlib", pyplot=plt)
    # numpy-1 paths used by dataloading/kitti360pose/*.py
    import numpy.lib as nplib

    if "numpy.lib.function_base" not in sys.modules:
        nplib.function_base = _module("numpy.lib.function_base", flip=np.flip)
    if "numpy.lib.arraysetops" not in sys.modules:
        nplib.arraysetops = _module("numpy.lib.arraysetops", isin=np.isin)
    gnn = _module(
        "torch_geometric.nn",
        fps=pyg_ops.fps,
        radius=pyg_ops.radius,
        PointConv=pyg_ops.PointConv,
        global_max_pool=pyg_ops.global_max_pool,
    )
    tr = _module(
        "torch_geometric.transforms",
        FixedPoints=pyg_ops.FixedPoints,
        NormalizeScale=pyg_ops.NormalizeScale,
        Compose=pyg_ops.Compose,
        RandomRotate=pyg_ops.RandomRotate,
    )
    data = _module("torch_geometric.data", Data=pyg_ops.Data, Batch=pyg_ops.Batch)
    _module("torch_geometric", nn=gnn, transforms=tr, data=data)

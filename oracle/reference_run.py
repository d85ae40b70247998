"""Run the reference's OWN Python for the coarse path.  TEST INFRASTRUCTURE, this container only.

Imports models/*, training/coarse.py::eval_epoch and evaluation/coarse.py::run_coarse
unchanged from /root/reference (never copied) under oracle/stubs.py, with the HF tokenizer/T5
replaced by oracle/fake_t5.py.  /root/reference does not exist on the GPU box, so nothing
under tests -m gpu, smoke() or bench.py calls this; it is used by oracle/make_golden.py and
by CPU tests that skip when the reference is absent.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

from . import fake_t5, stubs

REFERENCE_ROOT = os.environ.get("TEXT2LOC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "cell_retrieval.py"))


def default_args(**over) -> argparse.Namespace:
    """evaluation/args.py:7-89 defaults that define the hot-path shapes."""
    a = dict(
        batch_size=1, top_k=[1, 3, 5, 10], threshs=[5, 10, 15],
        use_features=["class", "color", "position", "num"],
        ranking_loss="pairwise", coarse_embed_dim=256, pointnet_layers=3, pointnet_variation=0,
        pointnet_numpoints=256, pointnet_path="", pointnet_freeze=False, pointnet_features=2,
        class_embed=False, color_embed=False, object_size=28,
        object_inter_module_num_heads=4, object_inter_module_num_layers=2,
        hungging_model="fake-t5", fixed_embedding=True,
        inter_module_num_heads=4, inter_module_num_layers=1,
        intra_module_num_heads=4, intra_module_num_layers=1,
    )
    a.update(over)
    return argparse.Namespace(**a)


_LOADED = {}


def load(fake_seed: int = 0):
    """Import the reference modules; returns a dict of the symbols the oracle uses."""
    if _LOADED:
        return _LOADED
    assert available(), f"reference not found at {REFERENCE_ROOT}"
    stubs.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import transformers

    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda name, *a, **k: fake_t5.FakeTokenizer())
    transformers.T5EncoderModel.from_pretrained = classmethod(lambda cls, name, *a, **k: fake_t5.FakeT5Encoder(fake_seed))
    with contextlib.redirect_stdout(io.StringIO()):
        from datapreparation.kitti360pose.imports import Cell, Object3d, Pose, DescriptionBestCell
        from datapreparation.kitti360pose.utils import COLOR_NAMES, KNOWN_CLASS
        from dataloading.kitti360pose.utils import batch_object_points
        from dataloading.kitti360pose.cells import Kitti360CoarseCellOnlyDataset, Kitti360CoarseDataset
        from models.cell_retrieval import CellRetrievalNetwork
        from training.coarse import eval_epoch
        from evaluation.coarse import run_coarse
    _LOADED.update(
        Cell=Cell, Object3d=Object3d, Pose=Pose, DescriptionBestCell=DescriptionBestCell,
        COLOR_NAMES=COLOR_NAMES, KNOWN_CLASS=KNOWN_CLASS, batch_object_points=batch_object_points,
        Kitti360CoarseCellOnlyDataset=Kitti360CoarseCellOnlyDataset,
        Kitti360CoarseDataset=Kitti360CoarseDataset,
        CellRetrievalNetwork=CellRetrievalNetwork, eval_epoch=eval_epoch, run_coarse=run_coarse,
    )
    return _LOADED


def fine_args(**over) -> argparse.Namespace:
    """evaluation/args.py defaults of the fine stage (:41-46, :80-81) on top of the shared ones."""
    return default_args(fine_embed_dim=128, fine_num_decoder_heads=4, fine_num_decoder_layers=2, pad_size=16, num_mentioned=6,
                        fine_intra_module_num_heads=4, fine_intra_module_num_layers=1, **over)


def build_fine_model(state_dict: dict, args=None, fake_seed: int = 0):
    """The reference CrossMatch (models/cross_matcher.py:39-129) carrying `state_dict`."""
    from synth import pointnet_state_dict

    ref = load(fake_seed)
    args = args or fine_args()
    with contextlib.redirect_stdout(io.StringIO()):
        from models.cross_matcher import CrossMatch
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
    with tempfile.TemporaryDirectory() as tmp:
        args.pointnet_path = os.path.join(tmp, "pointnet.pth")
        torch.save(pointnet_state_dict(sd), args.pointnet_path)
        with contextlib.redirect_stdout(io.StringIO()):
            model = CrossMatch(ref["KNOWN_CLASS"], ref["COLOR_NAMES"], args)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("llm_model" in k for k in res.missing_keys), res.missing_keys
    return model.eval()


def build_model(state_dict: dict, args=None, fake_seed: int = 0):
    """The reference CellRetrievalNetwork carrying `state_dict` (numpy or torch values)."""
    from synth import pointnet_state_dict

    ref = load(fake_seed)
    args = args or default_args()
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
    with tempfile.TemporaryDirectory() as tmp:
        args.pointnet_path = os.path.join(tmp, "pointnet.pth")
        torch.save(pointnet_state_dict(sd), args.pointnet_path)  # object_encoder.py:50 loads it unconditionally
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref["CellRetrievalNetwork"](ref["KNOWN_CLASS"], ref["COLOR_NAMES"], args)
    res = model.load_state_dict(sd, strict=False)  # evaluation/coarse.py:123
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("llm_model" in k for k in res.missing_keys), res.missing_keys
    return model.eval()

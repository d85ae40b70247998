"""Drop-in for models/cross_matcher.py::CrossMatch, the fine-localisation model (SURVEY.md section 8f row 1,
BASELINE configs[4]).

Same constructor signature, attributes and call as the reference class (models/cross_matcher.py:39-135):
``embed_dim``, ``device`` / ``get_device()``, ``eval()``, ``to(device)``, ``load_state_dict(sd, strict=False)`` and
``model(objects, hints, object_points) -> offsets [B, 2]``.  The arithmetic -- ObjectEncoder and
LanguageEncoder(is_fine) at d = 128, the cascaded cross-attention decoder layers, the offset MLP -- runs in the
sm_100a engine (t2l_fine_offsets); the frozen T5 stays in front, as for the coarse model.
"""
from __future__ import annotations

from typing import List

import torch

from . import dataio
from .engine import Engine, EngineError

_SUPPORTED = dict(
    fine_embed_dim=128, fine_num_decoder_heads=4, fine_num_decoder_layers=2, fine_intra_module_num_heads=4,
    fine_intra_module_num_layers=1, pointnet_layers=3, pointnet_variation=0, pointnet_numpoints=256, pointnet_features=2,
    class_embed=False, color_embed=False,
)


class CrossMatch:
    def __init__(self, known_classes: List[str], known_colors: List[str], args, text_frontend=None, device=None):
        for key, want in _SUPPORTED.items():
            got = getattr(args, key, want)
            if got != want:
                raise EngineError(f"args.{key}={got!r} is not supported by the B200 engine (built for {want!r}, the reference's eval defaults)")
        self.args = args
        self.embed_dim = args.fine_embed_dim
        self._engine = Engine(device)
        self._frontend = text_frontend
        self.training = False
        # opt-in: run_fine packs the retrieved cells' padded objects with the vectorised packer (dataio.pack_cell_database)
        self.vectorised_packing = False
        # run_fine: with a sentence-caching front end the hint encodings are cached per distinct (sentence, n_tok) too
        self.cache_sentence_rows = True
        from .text_frontend import SentenceRowCache

        self._sentence_rows = SentenceRowCache()

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise EngineError("the B200 engine is inference-only")
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise EngineError("the B200 engine has no CPU path")
        return self

    def load_state_dict(self, state_dict, strict: bool = False):
        """CrossMatch.state_dict() key names; llm_model.* keys are ignored (training/fine.py:274-279 never saves them)."""
        self._engine.load_state_dict({k: v for k, v in state_dict.items() if "llm_model" not in k})
        self._sentence_rows.clear()  # cached rows were computed with the previous weights
        return self

    @property
    def device(self):
        return self._engine.device

    def get_device(self):
        return self._engine.device

    @property
    def engine(self) -> Engine:
        return self._engine

    def frontend(self):
        if self._frontend is None:
            from .text_frontend import HFT5Frontend

            self._frontend = HFT5Frontend(self.args.hungging_model, self.device)
        return self._frontend

    @torch.no_grad()
    def encode_hints(self, descriptions: List[str]):
        """-> (hint encodings [len(descriptions) * n_hints, 128] on device, n_hints): LanguageEncoder(is_fine) behind the
        frozen T5 (models/language_encoder.py:108-140), one row per hint sentence.  Every row depends on its own sentence
        only, so a sentence-caching front end lets each distinct (sentence, n_tok) be encoded once."""
        fe = self.frontend()
        if self.cache_sentence_rows and hasattr(fe, "prepare") and hasattr(fe, "states"):
            sentences, n_hints, n_tok = fe.prepare(descriptions)
            return self._sentence_rows.rows(sentences, n_tok, lambda new: self._engine.fine_encode_hints(fe.states(new, n_tok))), n_hints
        feats, n_hints = fe(descriptions)
        return self._engine.fine_encode_hints(feats), n_hints

    @torch.no_grad()
    def forward(self, objects, hints: List[str], object_points) -> torch.Tensor:
        """objects: List[List[Object3d]] (every cell padded to pad_size), hints: one description per cell,
        object_points: one point batch per cell -> offsets FloatTensor [B, 2] on device (cross_matcher.py:83-129)."""
        if len(objects) != len(hints) or len(objects) != len(object_points):
            raise EngineError("CrossMatch.forward: objects, hints and object_points must have one entry per cell")
        pts, meta, cell_ptr = dataio.pack_cells(objects, object_points)
        feats, n_hints = self.frontend()(hints)
        return self._engine.fine_offsets(pts, meta, cell_ptr, feats, n_hints)

    __call__ = forward

"""Drop-in for models/cell_retrieval.py::CellRetrievalNetwork on the coarse-retrieval path.

Same constructor signature, attributes and methods as the reference class
(models/cell_retrieval.py:13-120): ``embed_dim``, ``object_size``, ``device`` /
``get_device()``, ``eval()``, ``to(device)``, ``load_state_dict(sd, strict=False)``,
``encode_text(descriptions)``, ``encode_objects(objects, object_points)``; ``forward()``
raises, as in the reference (:112-113).  The arithmetic runs in the sm_100a engine.

Inference only: the reference's eval path is under @torch.no_grad() (training/coarse.py:63);
training loops are out of scope.
"""
from __future__ import annotations

from typing import List

import torch

from . import dataio
from .engine import Engine, EngineError

_SUPPORTED = dict(
    coarse_embed_dim=256, pointnet_layers=3, pointnet_variation=0, pointnet_numpoints=256, pointnet_features=2,
    object_size=28, object_inter_module_num_heads=4, object_inter_module_num_layers=2,
    inter_module_num_heads=4, inter_module_num_layers=1, intra_module_num_heads=4, intra_module_num_layers=1,
    class_embed=False, color_embed=False,
)


class CellRetrievalNetwork:
    def __init__(self, known_classes: List[str], known_colors: List[str], args, text_frontend=None, device=None):
        """known_classes / known_colors are accepted for signature compatibility; they only size two
        classifier heads the path never uses (pointnet2.py:91-92).

        text_frontend: callable descriptions -> (t5 features [B*S, L, 1024], S), i.e. the frozen
        sentence-split + tokenizer + T5 encoder in front of the engine
        (models/language_encoder.py:108-125).  Defaults to text_frontend.HFT5Frontend(args.hungging_model).
        """
        for key, want in _SUPPORTED.items():
            got = getattr(args, key, want)
            if got != want:
                raise EngineError(f"args.{key}={got!r} is not supported by the B200 engine (built for {want!r}, the reference's eval defaults)")
        feats = list(getattr(args, "use_features", ["class", "color", "position", "num"]))
        if feats != ["class", "color", "position", "num"]:
            raise EngineError(f"use_features={feats} is not supported (engine is built for the 4-feature default)")
        self.args = args
        self.embed_dim = args.coarse_embed_dim
        self.object_size = args.object_size
        self._engine = Engine(device)
        self._frontend = text_frontend
        self.training = False
        # opt-in: eval_epoch packs the cell database once (vectorised) and re-uses it across evaluations (evaluation.py)
        self.cache_packed_cells = False
        # with a front end that caches T5 states per sentence (text_frontend.SentenceCacheFrontend) the token stage of the
        # text head is cached per distinct sentence as well; set False to push every batch through it whole
        self.cache_sentence_rows = True
        from .text_frontend import SentenceRowCache

        self._sentence_rows = SentenceRowCache()

    # ---- nn.Module surface the eval drivers touch -----------------------------------------
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
        if torch.device(device).index not in (None, self._engine.device.index):
            raise EngineError(f"engine was created on {self._engine.device}; create a new CellRetrievalNetwork for {device}")
        return self

    def load_state_dict(self, state_dict, strict: bool = False):
        """Reference key names; llm_model.* keys are ignored as the reference
        checkpoints omit them (training/coarse.py:329-331)."""
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

    # ---- the two encoders -----------------------------------------------------------------
    @torch.no_grad()
    def encode_text(self, descriptions: List[str]) -> torch.Tensor:
        """[B] strings -> FloatTensor [B, 256], unit rows, on device (cell_retrieval.py:57-63)."""
        if self._frontend is None:
            from .text_frontend import HFT5Frontend

            self._frontend = HFT5Frontend(self.args.hungging_model, self.device)
        fe = self._frontend
        if self.cache_sentence_rows and hasattr(fe, "prepare") and hasattr(fe, "states"):
            # a caching front end knows the batch's sentences: the token stage runs once per distinct (sentence, n_tok)
            # and the batch is assembled from the cached rows (text_frontend.SentenceRowCache)
            sentences, n_sent, n_tok = fe.prepare(descriptions)
            pooled = self._sentence_rows.rows(sentences, n_tok, lambda new: self._engine.encode_text_tokens(fe.states(new, n_tok)))
            return self._engine.encode_text_sentences(pooled, n_sent)
        feats, n_sent = fe(descriptions)
        return self._engine.encode_text(feats, n_sent)

    @torch.no_grad()
    def encode_text_features(self, t5_features: torch.Tensor, n_sent: int) -> torch.Tensor:
        """Engine-level entry: T5 last_hidden_state [B*S, L, 1024] -> [B, 256]."""
        return self._engine.encode_text(t5_features, n_sent)

    @torch.no_grad()
    def encode_objects(self, objects, object_points) -> torch.Tensor:
        """objects: List[List[Object3d]], object_points: List[Batch] (one point batch per cell)
        -> FloatTensor [B, 256], unit rows, on device (cell_retrieval.py:65-110)."""
        pts, meta, cell_ptr = dataio.pack_cells(objects, object_points)
        return self._engine.encode_cells(pts, meta, cell_ptr)

    @torch.no_grad()
    def encode_cells_packed(self, pts, meta, cell_ptr) -> torch.Tensor:
        """Engine-level entry on the packed layout."""
        return self._engine.encode_cells(pts, meta, cell_ptr)

    def forward(self):
        raise Exception("Not implemented.")

    __call__ = forward

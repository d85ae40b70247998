"""The frozen front of the language branch: sentence split -> tokenizer -> T5 encoder.

models/language_encoder.py:108-125.  north_star scopes the engine to "the language_encoder over
frozen T5 text embeddings": T5 (a third-party pretrained model, ~27x the text head's FLOPs) is an
INPUT producer here, run with the stock HF implementation, and the engine starts at
``last_hidden_state``.  No T5 weights exist offline, so tests inject a deterministic fake with
the same call signature.
"""
from __future__ import annotations

import re
from typing import List, Tuple

import torch

_SENT_END = re.compile(r"(?<=[.!?])\s+")
_WARNED = False


def split_sentences(text: str) -> List[str]:
    """nltk.tokenize.sent_tokenize, as the reference (models/language_encoder.py:110).  Only when nltk itself is not
    installed does a regex split stand in -- exact for the reference's templated hints
    (dataloading/kitti360pose/base.py:60-68), announced once because sentence boundaries of free text may differ.
    An installed nltk without its punkt model raises, with the fix in the message: a silent fallback would change
    n_sent and with it the embeddings."""
    global _WARNED
    try:
        from nltk import tokenize
    except ImportError:
        if not _WARNED:
            import warnings

            warnings.warn("nltk is not installed: sentences are split on [.!?] + whitespace, which matches nltk only for "
                          "templated hint text; install nltk and its punkt model for the reference's behaviour")
            _WARNED = True
        return [s for s in _SENT_END.split(text.strip()) if s]
    try:
        return tokenize.sent_tokenize(text)
    except LookupError as err:
        raise LookupError("nltk is installed but its punkt model is missing: run `python -m nltk.downloader punkt` "
                          "(the reference hard-depends on it, models/language_encoder.py:110)") from err


class HFT5Frontend:
    def __init__(self, model_name: str, device):
        from transformers import AutoTokenizer, T5EncoderModel

        self.tokenizer = AutoTokenizer.from_pretrained(model_name)
        self.model = T5EncoderModel.from_pretrained(model_name).to(device).eval()
        self.device = device

    @torch.no_grad()
    def __call__(self, descriptions: List[str]) -> Tuple[torch.Tensor, int]:
        sentences: List[str] = []
        for d in descriptions:
            sentences.extend(split_sentences(d))
        n_sent = len(sentences) // len(descriptions)  # the reference assumes equal counts (:114)
        tok = self.tokenizer(sentences, return_tensors="pt", padding="longest")
        out = self.model(input_ids=tok["input_ids"].to(self.device), attention_mask=tok["attention_mask"].to(self.device),
                         output_attentions=False)
        return out.last_hidden_state.float().contiguous(), n_sent

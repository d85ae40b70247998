"""Deterministic stand-in for the HF tokenizer + frozen T5 encoder.  TEST INFRASTRUCTURE.

The T5 encoder is an *input* to the path (north_star: "language_encoder over frozen T5 text
embeddings"; the engine starts at last_hidden_state, models/language_encoder.py:125), and no
T5 weights exist offline.  Both the reference run and the drop-in's text front-end use this
fake so they see identical [B*S, L, 1024] features.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch
import torch.nn as nn

VOCAB = 512
DIM = 1024


class FakeTokenizer:
    pad_id = 0

    def __call__(self, sentences, return_tensors="pt", padding="longest"):
        ids = [[1 + zlib.crc32(w.encode()) % (VOCAB - 1) for w in s.replace(".", " .").split()] for s in sentences]
        n = max(len(i) for i in ids)
        input_ids = torch.tensor([i + [self.pad_id] * (n - len(i)) for i in ids], dtype=torch.long)
        return {"input_ids": input_ids, "attention_mask": (input_ids != self.pad_id).long()}


class _Out:
    def __init__(self, h):
        self.last_hidden_state = h


class FakeT5Encoder(nn.Module):
    """embedding lookup + a fixed positional term, scale ~0.2 like real T5 encoder outputs."""

    def __init__(self, seed: int = 0):
        super().__init__()
        rng = np.random.default_rng([seed, 0x7F5])
        self.encoder = nn.Module()
        self.encoder.embed_tokens = nn.Embedding(VOCAB, DIM)
        with torch.no_grad():
            self.encoder.embed_tokens.weight.copy_(torch.from_numpy((rng.standard_normal((VOCAB, DIM)) * 0.2).astype(np.float32)))
        self.register_buffer("pos", torch.from_numpy((rng.standard_normal((64, DIM)) * 0.05).astype(np.float32)))

    def forward(self, input_ids=None, attention_mask=None, output_attentions=False):
        h = self.encoder.embed_tokens(input_ids) + self.pos[: input_ids.shape[1]][None]
        return _Out(h)


class FakeFrontend:
    """descriptions -> (t5 features f32 [B*S, L, 1024], S).  Same steps as
    models/language_encoder.py:108-125 with the fakes above."""

    def __init__(self, seed: int = 0, sent_tokenize=None):
        from .stubs import sent_tokenize as st

        self.sent_tokenize = sent_tokenize or st
        self.tokenizer = FakeTokenizer()
        self.model = FakeT5Encoder(seed).eval()

    @torch.no_grad()
    def __call__(self, descriptions):
        sents = []
        for d in descriptions:
            sents.extend(self.sent_tokenize(d))
        n_sent = len(sents) // len(descriptions)
        tok = self.tokenizer(sents)
        h = self.model(input_ids=tok["input_ids"], attention_mask=tok["attention_mask"]).last_hidden_state
        return h.contiguous(), n_sent

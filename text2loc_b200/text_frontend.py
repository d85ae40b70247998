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


class SentenceCacheFrontend:
    """Sentence-level cache in front of the frozen encoder.

    The reference pushes every description's sentences through T5-large on every call (models/language_encoder.py:116-125),
    ~27x the FLOPs of the text head behind it, although the templated hint vocabulary is tiny (direction x colour x class =
    5 x 8 x 22 sentences, dataloading/kitti360pose/base.py:60-68).  Each distinct sentence is encoded ONCE, padded to `cap`
    tokens, and its states [cap, 1024] are kept on the device; a batch is assembled by slicing every sentence's states to the
    batch's longest token count (the reference pads with padding="longest").  This reproduces the uncached output because an
    encoder with a key-padding mask never lets a position attend to pads: the state at position p -- pad positions included,
    which the reference feeds UNMASKED into its intra-module -- depends only on the sentence's real
    tokens and on p (T5's relative position buckets), not on how many pads follow.

    tokenizer(sentences, return_tensors="pt", padding="longest") -> {"input_ids", "attention_mask"};
    model(input_ids=, attention_mask=).last_hidden_state."""

    def __init__(self, tokenizer, model, device, cap: int = 32, pad_id: int = 0, split=None):
        self.tokenizer, self.model, self.device, self.cap, self.pad_id = tokenizer, model, device, cap, pad_id
        self.split = split or split_sentences
        self.cache = {}  # sentence -> (n_tokens, states [cap, 1024] on device)
        self.encoder_calls = 0

    @torch.no_grad()
    def _encode_new(self, sentences: List[str]):
        tok = self.tokenizer(sentences, return_tensors="pt", padding="longest")
        ids, mask = tok["input_ids"], tok["attention_mask"]
        n_tok = mask.sum(dim=1)
        if ids.shape[1] > self.cap:
            raise ValueError(f"a sentence has {ids.shape[1]} tokens; raise SentenceCacheFrontend(cap=...) (the engine takes up to 32)")
        pad = self.cap - ids.shape[1]
        if pad:
            ids = torch.cat([ids, torch.full((ids.shape[0], pad), self.pad_id, dtype=ids.dtype)], dim=1)
            mask = torch.cat([mask, torch.zeros((mask.shape[0], pad), dtype=mask.dtype)], dim=1)
        out = self.model(input_ids=ids.to(self.device), attention_mask=mask.to(self.device), output_attentions=False).last_hidden_state
        self.encoder_calls += 1
        for s, n, h in zip(sentences, n_tok.tolist(), out.float()):
            self.cache[s] = (int(n), h.contiguous())

    @torch.no_grad()
    def prepare(self, descriptions: List[str]) -> Tuple[List[str], int, int]:
        """-> (the batch's sentences in order, sentences per description, the batch's longest token count); every
        sentence is in the cache afterwards."""
        sentences: List[str] = []
        for d in descriptions:
            sentences.extend(self.split(d))
        n_sent = len(sentences) // len(descriptions)  # the reference assumes equal counts (:114)
        new = [s for s in dict.fromkeys(sentences) if s not in self.cache]
        if new:
            self._encode_new(new)
        return sentences, n_sent, max(self.cache[s][0] for s in sentences)  # padding="longest" over THIS batch

    def states(self, sentences: List[str], n_tok: int) -> torch.Tensor:
        """Cached states of `sentences`, each cut to n_tok positions -> [len(sentences), n_tok, 1024]."""
        return torch.stack([self.cache[s][1][:n_tok] for s in sentences]).contiguous()

    @torch.no_grad()
    def __call__(self, descriptions: List[str]) -> Tuple[torch.Tensor, int]:
        sentences, n_sent, n_tok = self.prepare(descriptions)
        return self.states(sentences, n_tok), n_sent


class SentenceRowCache:
    """Per-sentence rows of the text head behind a SentenceCacheFrontend, computed once per distinct (sentence, n_tok).

    Everything the language encoder does BEFORE its sentence-level module is a function of one sentence's token states
    alone: intra_module over the sentence's tokens and the max over them (models/language_encoder.py:130-133; for the fine
    model also inter_mlp, :137-140).  That token stage is 99 % of the text head's arithmetic (906 of
    914 MMAC per query), and with the templated hint vocabulary a batch of B x 6 sentences holds only a few hundred distinct
    ones.  The cache keeps one row per (sentence, n_tok) -- n_tok is part of the key because the reference feeds pad
    positions UNMASKED into intra_module, so the row depends on how far the batch was padded -- in one device table and
    assembles a batch with a single index_select.  The engine's token stage is row-independent bit for bit
    (tests/test_gpu_parity.py::test_encode_text_host_streaming_equals_device), so the result equals the uncached call."""

    def __init__(self, max_rows: int = 1 << 16):
        self.max_rows = max_rows
        self.index = {}    # (sentence, n_tok) -> row of the table
        self.table = None  # [rows, width] on the engine's device
        self.computed = 0  # rows ever computed (tests / diagnostics)

    def clear(self):
        """Forget every row (the weights behind `compute` changed)."""
        self.index, self.table = {}, None

    def rows(self, sentences: List[str], n_tok: int, compute) -> torch.Tensor:
        """compute(list of distinct missing sentences) -> tensor [len, width]; returns [len(sentences), width]."""
        keys = [(s, n_tok) for s in sentences]
        missing = [k for k in dict.fromkeys(keys) if k not in self.index]
        if missing and len(self.index) + len(missing) > self.max_rows:  # bounded: start over rather than grow without limit
            self.clear()
            missing = list(dict.fromkeys(keys))
        if missing:
            new = compute([s for s, _ in missing])
            base = 0 if self.table is None else self.table.shape[0]
            self.table = new.clone() if self.table is None else torch.cat([self.table, new])
            self.index.update({k: base + i for i, k in enumerate(missing)})
            self.computed += len(missing)
        sel = torch.tensor([self.index[k] for k in keys], dtype=torch.long, device=self.table.device)
        return self.table.index_select(0, sel)


class HFT5Frontend:
    def __init__(self, model_name: str, device):
        from transformers import AutoTokenizer, T5EncoderModel

        self.tokenizer = AutoTokenizer.from_pretrained(model_name)
        self.model = T5EncoderModel.from_pretrained(model_name).to(device).eval()
        self.device = device

    def cached(self, cap: int = 32) -> SentenceCacheFrontend:
        """The same front end with the sentence-level cache (each distinct sentence goes through T5 once)."""
        return SentenceCacheFrontend(self.tokenizer, self.model, self.device, cap, pad_id=self.tokenizer.pad_token_id or 0)

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

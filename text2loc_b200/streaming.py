"""Streamed encode + score for databases that do not fit in HBM at once (BASELINE configs[3]: 1M cells x 16 objects).

The reference encodes every cell into one big array and then scores it (training/coarse.py:105-125).  Here the database
is walked in chunks: encode a chunk of cells -> register it as the engine's shard with its global row offset -> search ->
fold the chunk's top-k into a running per-query top-k (t2l_search_topk_accumulate).  Neither the points nor the full
[N, 256] embedding matrix are ever resident; the result equals the unstreamed search bit for bit because every list is
ordered by the same (fp64 score desc, row asc) relation.
"""
from __future__ import annotations

from typing import Callable, Iterable, Tuple

import numpy as np
import torch


class StreamingRetrieval:
    """Running top-k over chunks of cells.  `add_cells` takes the packed layout."""

    def __init__(self, engine, queries: torch.Tensor, k: int):
        self.engine = engine
        self.Q = queries
        self.k = k
        self.idx, self.score = engine.new_running_topk(queries.shape[0], k)
        self.n_rows = 0
        self.n_fallback = torch.zeros(1, dtype=torch.int64, device=engine.device)

    def add_cells(self, pts, meta, cell_ptr, row_offset: int = None) -> torch.Tensor:
        """Encode one chunk of cells and fold it in; returns the chunk's embeddings (device)."""
        D = self.engine.encode_cells(pts, meta, cell_ptr)
        self.add_embeddings(D, row_offset)
        return D

    def add_embeddings(self, D: torch.Tensor, row_offset: int = None):
        off = self.n_rows if row_offset is None else int(row_offset)
        self.engine.db_build(D, row_offset=off)
        self.n_fallback += self.engine.search_topk_accumulate(self.Q, self.k, self.idx, self.score)
        self.n_rows = max(self.n_rows, off + D.shape[0])

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.idx, self.score


def stream_synthetic(engine, queries: torch.Tensor, k: int, seed: int, first_cell: int, n_cells: int, obj_per_cell: int,
                     chunk_cells: int = 1024, keep_embeddings: bool = False):
    """configs[3] driver: cells [first_cell, first_cell + n_cells) are generated on the device chunk by chunk
    (t2l_synth_cells), encoded and scored.  Returns (idx, score, n_fallback[, embeddings])."""
    sr = StreamingRetrieval(engine, queries, k)
    bufs = None
    kept = []
    for c0 in range(first_cell, first_cell + n_cells, chunk_cells):
        nc = min(chunk_cells, first_cell + n_cells - c0)
        if bufs is None:
            n = chunk_cells * obj_per_cell
            bufs = (torch.empty((n, 256, 6), dtype=torch.float32, device=engine.device), torch.empty((n, 7), dtype=torch.float32, device=engine.device))
        pts, meta, ptr = engine.synth_cells(seed, c0, nc, obj_per_cell, out=bufs)
        D = sr.add_cells(pts, meta, ptr, row_offset=c0)
        if keep_embeddings:
            kept.append(D)
    idx, score = sr.result()
    if keep_embeddings:
        return idx, score, sr.n_fallback, torch.cat(kept)
    return idx, score, sr.n_fallback

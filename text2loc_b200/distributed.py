"""Row-sharded database search across the GPUs of one box.

The reference is single-process; the only natural parallel axis of the path is the database:
rank g owns cell rows [g*N/G, (g+1)*N/G), encodes and scores them, and ONE exchange step joins
the per-shard top-k lists: an all-gather of [nq, k] x (f64 score, i64 row) followed by the
merge kernel (t2l_merge_topk).  The result is independent of G by construction: every list is
ordered by the same (score desc, row asc) relation on the same fp64 scores.

torch.distributed is plumbing only (NCCL over NVLink on GPUs; gloo in the CPU tests, where the
merge itself is checked with the numpy oracle).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`; sizes differ by at most one row."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """[m, ...] per rank (same shape on every rank) -> [G, m, ...]."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    if ws == 1:
        return x.unsqueeze(0)
    out = torch.empty((ws,) + tuple(x.shape), dtype=x.dtype, device=x.device)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)  # one NCCL all-gather over NVLink
    else:
        dist.all_gather(list(out.unbind(0)), x.contiguous(), group=group)  # gloo (CPU tests)
    return out


def merge_topk_host(idx_all: np.ndarray, score_all: np.ndarray, k: int):
    """Reference merge on the host (numpy): same order relation as the device kernel.
    idx_all / score_all: [G, nq, k]; empty slots carry idx -1."""
    G, nq, kk = idx_all.shape
    idx = np.transpose(idx_all, (1, 0, 2)).reshape(nq, G * kk)
    sc = np.transpose(score_all, (1, 0, 2)).reshape(nq, G * kk).copy()
    sc[idx < 0] = -np.inf
    big = np.where(idx < 0, np.iinfo(np.int64).max, idx)
    order = np.lexsort((big, -sc), axis=1)[:, :k]  # primary: score desc, secondary: row asc
    return np.take_along_axis(idx, order, 1), np.take_along_axis(sc, order, 1)


def sharded_search(engine, Q_local_or_all: torch.Tensor, k: int, queries_are_sharded: bool = False, group=None, timers=None):
    """Search this rank's shard (engine.db_build must hold it, with its row_offset) for ALL queries
    and merge across ranks.  Returns (idx [nq,k], score [nq,k], n_fallback) on every rank.

    queries_are_sharded: each rank holds a different slice of the query embeddings (the text head
    was split by queries); they are all-gathered first.

    ONE exchange step joins the results: the rank's (idx, score) pair lives in one packed buffer of 8-byte words
    ([2, nq, k]: row indices, then the fp64 scores' bits), so a single all-gather moves both, and t2l_merge_topk_packed
    reads the gathered [G, 2, nq, k] block in place.
    timers: optional dict of lists; when given, CUDA events bracket each stage and their milliseconds are appended under
    'allgather_q', 'search', 'allgather_topk', 'merge' (this synchronises the device: measurement passes only).
    """
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    marks = []

    def mark(name):
        if timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    Q = Q_local_or_all
    if queries_are_sharded:
        Q = all_gather_rows(Q, group).reshape(-1, Q.shape[-1])
    mark("allgather_q")
    nq = Q.shape[0]
    if ws == 1:
        idx, score, nfb = engine.search_topk(Q, k)
        mark("search")
    else:
        packed = torch.empty((2, nq, k), dtype=torch.int64, device=Q.device)
        _, _, nfb = engine.search_topk(Q, k, out=(packed[0], packed[1].view(torch.float64)))
        mark("search")
        gathered = all_gather_rows(packed, group)  # [G, 2, nq, k]
        mark("allgather_topk")
        idx, score = engine.merge_topk_packed(gathered, nq, k)
        mark("merge")
    if timers is not None:
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            timers.setdefault(name, []).append(a.elapsed_time(b))
    return idx, score, nfb

"""Thin torch-facing wrapper over the C ABI: tensors in, tensors out, raw pointers across.

PyTorch is used for device memory and streams only; all arithmetic happens in
libtext2loc_b200.so.  Work is enqueued on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib, weights

EMBED_DIM = 256
T5_DIM = 1024
NUM_POINTS = 256
MAX_TOPK_FAST = 12


class EngineError(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def host_chunk_schedule(nq: int, cq: int):
    """Queries per H2D staging chunk of a host-streamed encode_text: full chunks of `cq` queries, with a quarter and a half
    chunk at both ends when the batch is long enough (the first copy and the last chunk's compute are what stays exposed).
    Every entry is in 1..cq and the entries sum to nq."""
    if nq <= 0:
        return []
    if nq >= 6 * cq and cq >= 4:
        head = [cq // 4, cq // 2]
        body = nq - 2 * sum(head)
        return head + [cq] * (body // cq) + ([body % cq] if body % cq else []) + head[::-1]
    return [cq] * (nq // cq) + ([nq % cq] if nq % cq else [])


class Engine:
    """One engine per CUDA device (t2l_create / t2l_destroy)."""

    def __init__(self, device=None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise EngineError("text2loc_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise EngineError(f"text2loc_b200 runs on CUDA devices only, got {self.device}")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        h = ctypes.c_void_p()
        if self._lib.t2l_create(index, ctypes.byref(h)) != 0:
            raise EngineError(self._lib.t2l_last_error(None).decode())
        self._h = h
        self._db = None  # keeps the database tensor alive (the engine holds a raw pointer to it)
        self._stage = None  # device staging pair of the host-streaming encode_text
        self._copy_stream = None
        self._stage_done = [None, None]
        self.has_weights = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.t2l_destroy(h)
            self._h = None

    # ---- plumbing -------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise EngineError(self._lib.t2l_last_error(self._h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t, dtype):
        t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype).contiguous()
        return t

    @property
    def launch_count(self) -> int:
        return int(self._lib.t2l_launch_count(self._h))

    # ---- weights --------------------------------------------------------------------------
    def load_state_dict(self, state_dict: dict):
        """Reference checkpoint keys -> folded engine weights (text2loc_b200/weights.py)."""
        for name, arr in weights.engine_weights(state_dict).items():
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            self._check(self._lib.t2l_set_weight(
                self._h, name.encode(), arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), arr.shape[0], arr.shape[1]))
        self._check(self._lib.t2l_finalize_weights(self._h))
        self.has_weights = True

    def reserve(self, max_objects: int = 0, max_cells: int = 0, max_sentences: int = 0, max_tokens: int = 0, max_queries: int = 0):
        """Size the workspace up front (t2l_reserve) so that steady-state calls never re-allocate."""
        self._check(self._lib.t2l_reserve(self._h, max_objects, max_cells, max_sentences, max_tokens, max_queries))

    # ---- encoders -------------------------------------------------------------------------
    def encode_cells(self, pts, meta, cell_ptr) -> torch.Tensor:
        """pts [n,256,6], meta [n,7], cell_ptr int32 [B+1] (host) -> unit rows [B,256] on device."""
        pts = self._dev(pts, torch.float32)
        meta = self._dev(meta, torch.float32)
        cp = np.ascontiguousarray(np.asarray(cell_ptr.cpu() if torch.is_tensor(cell_ptr) else cell_ptr), dtype=np.int32)
        n_cells = len(cp) - 1
        if pts.dim() != 3 or pts.shape[1:] != (NUM_POINTS, 6) or meta.shape != (pts.shape[0], 7) or cp[-1] != pts.shape[0]:
            raise EngineError(f"encode_cells: bad shapes pts {tuple(pts.shape)} meta {tuple(meta.shape)} cell_ptr[-1]={cp[-1]}")
        out = torch.empty((n_cells, EMBED_DIM), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_encode_cells(
            self._h, _ptr(pts), _ptr(meta), cp.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n_cells, _ptr(out), self._stream()))
        return out

    def encode_objects_debug(self, pts, cell_ptr) -> dict:
        """PointNet++ features2 and the FPS / ball-query index sets (parity tests)."""
        pts = self._dev(pts, torch.float32)
        cp = np.ascontiguousarray(np.asarray(cell_ptr.cpu() if torch.is_tensor(cell_ptr) else cell_ptr), dtype=np.int32)
        n = pts.shape[0]
        u8 = lambda *s: torch.zeros(s, dtype=torch.uint8, device=self.device)
        r = dict(features2=torch.empty((n, 256), dtype=torch.float32, device=self.device),
                 fps1=u8(n, 128), fps2=u8(n, 64), fps3=u8(n, 32), nbr1=u8(n, 128, 32), nbr2=u8(n, 64, 32), nbr3=u8(n, 32, 32),
                 cnt1=u8(n, 128), cnt2=u8(n, 64), cnt3=u8(n, 32))
        self._check(self._lib.t2l_encode_objects_debug(
            self._h, _ptr(pts), cp.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(cp) - 1,
            *[_ptr(r[k]) for k in ("features2", "fps1", "fps2", "fps3", "nbr1", "nbr2", "nbr3", "cnt1", "cnt2", "cnt3")], self._stream()))
        return r

    # tokens per H2D staging chunk of a HOST input (the copy of chunk i+1 hides behind the compute of chunk i)
    TOKENS_PER_CHUNK = int(os.environ.get("T2L_HOST_TOK_CHUNK", "16384"))

    def encode_text(self, t5, n_sent: int) -> torch.Tensor:
        """t5 [nq*n_sent, n_tok, 1024] (T5 last_hidden_state, float32 or float16) -> unit rows [nq,256] on device.

        float16 features take the fp16 entry point (t2l_encode_text_tokens_f16): half the bytes to move, no conversion
        kernel; the token layer computes on fp16 operand copies either way.
        A HOST tensor (ideally pinned) is streamed: the H2D copy of chunk i+1 runs on a side
        stream while the engine works on chunk i, so PCIe time hides behind compute."""
        t5 = torch.as_tensor(t5)
        if t5.dim() != 3 or t5.shape[2] != T5_DIM or t5.shape[0] % n_sent or t5.dtype not in (torch.float32, torch.float16):
            raise EngineError(f"encode_text: need float32 / float16 [nq*{n_sent}, n_tok, 1024], got {t5.dtype} {tuple(t5.shape)}")
        half = t5.dtype == torch.float16
        tokens = self._lib.t2l_encode_text_tokens_f16 if half else self._lib.t2l_encode_text_tokens
        nq, n_tok = t5.shape[0] // n_sent, t5.shape[1]
        out = torch.empty((nq, EMBED_DIM), dtype=torch.float32, device=self.device)
        if t5.is_cuda:
            t5 = self._dev(t5, t5.dtype)
            if not half:
                self._encode_text_dev(t5, n_sent, out)
                return out
            pooled = torch.empty((nq * n_sent, T5_DIM), dtype=torch.float32, device=self.device)
            self._check(tokens(self._h, _ptr(t5), nq * n_sent, n_tok, _ptr(pooled), self._stream()))
            self._check(self._lib.t2l_encode_text_sentences(self._h, _ptr(pooled), nq, n_sent, _ptr(out), self._stream()))
            return out
        t5 = t5.contiguous()
        cq = max(1, self.TOKENS_PER_CHUNK // (n_sent * n_tok))
        rows_per_chunk = cq * n_sent
        cur = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)  # one copy stream for the engine's lifetime
        if (self._stage is None or self._stage[0].shape[0] < rows_per_chunk or self._stage[0].shape[1] != n_tok
                or self._stage[0].dtype != t5.dtype):
            # (Re)allocation happens on the current stream: the caching allocator may hand back blocks that kernels
            # already queued on this stream still read (the previous staging pair included).  The copy stream must
            # not write them before that work has finished, and the blocks must not be recycled under the copies.
            self._stage = [torch.empty((rows_per_chunk, n_tok, T5_DIM), dtype=t5.dtype, device=self.device) for _ in range(2)]
            for t in self._stage:
                t.record_stream(self._copy_stream)
            self._copy_stream.wait_stream(cur)
            self._stage_done = [None, None]
        pooled = torch.empty((nq * n_sent, T5_DIM), dtype=torch.float32, device=self.device)
        # Chunk schedule: the copy of chunk i+1 runs under the compute of chunk i, so what stays exposed is the FIRST copy and
        # the LAST chunk's compute -- both start and end with quarter and half chunks.  The sentence stage runs per group of
        # ~4 chunks instead of once at the end (it would otherwise sit, whole, behind the last copy).
        sizes = host_chunk_schedule(nq, cq)
        group_q = 4 * cq
        q0 = g0 = 0
        for i, n_q in enumerate(sizes):
            q1 = q0 + n_q
            b = i % 2
            n_rows = n_q * n_sent
            with torch.cuda.stream(self._copy_stream):
                if self._stage_done[b] is not None:
                    self._copy_stream.wait_event(self._stage_done[b])  # the engine is done reading this buffer
                self._stage[b][:n_rows].copy_(t5[q0 * n_sent:q1 * n_sent], non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(self._copy_stream)
            cur.wait_event(copied)
            self._check(tokens(  # token stage of this chunk
                self._h, _ptr(self._stage[b]), n_rows, n_tok, _ptr(pooled[q0 * n_sent:q1 * n_sent]), self._stream()))
            self._stage_done[b] = torch.cuda.Event()
            self._stage_done[b].record(cur)
            if q1 - g0 >= group_q or q1 == nq:  # sentence stage of the queries whose tokens are done
                self._check(self._lib.t2l_encode_text_sentences(self._h, _ptr(pooled[g0 * n_sent:q1 * n_sent]), q1 - g0, n_sent,
                                                                _ptr(out[g0:q1]), self._stream()))
                g0 = q1
            q0 = q1
        return out

    def encode_text_tokens(self, t5) -> torch.Tensor:
        """Token stage alone (t2l_encode_text_tokens / _f16): device t5 [n_sentences, n_tok, 1024] float32 or float16 ->
        [n_sentences, 1024] = intra_module + max over tokens (models/language_encoder.py:130-133).  Rows are independent."""
        t5 = torch.as_tensor(t5)
        if t5.dim() != 3 or t5.shape[2] != T5_DIM or t5.dtype not in (torch.float32, torch.float16):
            raise EngineError(f"encode_text_tokens: need float32 / float16 [n_sentences, n_tok, 1024], got {t5.dtype} {tuple(t5.shape)}")
        t5 = self._dev(t5, t5.dtype)
        tokens = self._lib.t2l_encode_text_tokens_f16 if t5.dtype == torch.float16 else self._lib.t2l_encode_text_tokens
        pooled = torch.empty((t5.shape[0], T5_DIM), dtype=torch.float32, device=self.device)
        self._check(tokens(self._h, _ptr(t5), t5.shape[0], t5.shape[1], _ptr(pooled), self._stream()))
        return pooled

    def encode_text_sentences(self, pooled, n_sent: int) -> torch.Tensor:
        """Sentence stage alone (t2l_encode_text_sentences): pooled [nq*n_sent, 1024] -> unit rows [nq, 256]
        (inter_mlp, inter_module with `x += layer(x)`, max over sentences, normalise; language_encoder.py:137-148)."""
        pooled = self._dev(pooled, torch.float32)
        if pooled.dim() != 2 or pooled.shape[1] != T5_DIM or pooled.shape[0] % n_sent:
            raise EngineError(f"encode_text_sentences: need [nq*{n_sent}, 1024], got {tuple(pooled.shape)}")
        nq = pooled.shape[0] // n_sent
        out = torch.empty((nq, EMBED_DIM), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_encode_text_sentences(self._h, _ptr(pooled), nq, n_sent, _ptr(out), self._stream()))
        return out

    def _encode_text_dev(self, t5: torch.Tensor, n_sent: int, out: torch.Tensor):
        nq = t5.shape[0] // n_sent
        self._check(self._lib.t2l_encode_text(self._h, _ptr(t5), nq, n_sent, t5.shape[1], _ptr(out), self._stream()))

    # ---- fine stage -----------------------------------------------------------------------
    FINE_DIM = 128

    def fine_offsets(self, pts, meta, cell_ptr, t5, n_hints: int) -> torch.Tensor:
        """CrossMatch.forward on packed inputs: one (cell, description) pair per cell -> offsets [n_cells, 2]."""
        pts, meta = self._dev(pts, torch.float32), self._dev(meta, torch.float32)
        t5 = self._dev(t5, torch.float32)
        cp = np.ascontiguousarray(np.asarray(cell_ptr.cpu() if torch.is_tensor(cell_ptr) else cell_ptr), dtype=np.int32)
        n_cells = len(cp) - 1
        if t5.dim() != 3 or t5.shape[0] != n_cells * n_hints or t5.shape[2] != T5_DIM or cp[-1] != pts.shape[0]:
            raise EngineError(f"fine_offsets: t5 {tuple(t5.shape)} for {n_cells} cells x {n_hints} hints, pts {tuple(pts.shape)}")
        out = torch.empty((n_cells, 2), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_fine_offsets(self._h, _ptr(pts), _ptr(meta), cp.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n_cells,
                                               _ptr(t5), n_hints, t5.shape[1], _ptr(out), self._stream()))
        return out

    def fine_encode_objects(self, pts, meta, cell_ptr) -> torch.Tensor:
        pts, meta = self._dev(pts, torch.float32), self._dev(meta, torch.float32)
        cp = np.ascontiguousarray(np.asarray(cell_ptr.cpu() if torch.is_tensor(cell_ptr) else cell_ptr), dtype=np.int32)
        out = torch.empty((pts.shape[0], self.FINE_DIM), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_fine_encode_objects(self._h, _ptr(pts), _ptr(meta), cp.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                      len(cp) - 1, _ptr(out), self._stream()))
        return out

    def fine_encode_hints(self, t5) -> torch.Tensor:
        t5 = self._dev(t5, torch.float32)
        out = torch.empty((t5.shape[0], self.FINE_DIM), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_fine_encode_hints(self._h, _ptr(t5), t5.shape[0], t5.shape[1], _ptr(out), self._stream()))
        return out

    def fine_match(self, obj_emb, pair_cell, hints, pair_query, n_obj: int, n_hints: int) -> torch.Tensor:
        obj_emb, hints = self._dev(obj_emb, torch.float32), self._dev(hints, torch.float32)
        pair_cell = self._dev(pair_cell, torch.int32) if pair_cell is not None else None
        pair_query = self._dev(pair_query, torch.int32) if pair_query is not None else None
        n_pairs = len(pair_cell) if pair_cell is not None else (len(pair_query) if pair_query is not None else obj_emb.shape[0] // n_obj)
        out = torch.empty((n_pairs, 2), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_fine_match(self._h, _ptr(obj_emb), _ptr(pair_cell), _ptr(hints), _ptr(pair_query), n_pairs, n_obj, n_hints,
                                             _ptr(out), self._stream()))
        return out

    # ---- search ---------------------------------------------------------------------------
    def db_build(self, D, row_offset: int = 0):
        D = self._dev(D, torch.float32)
        if D.dim() != 2 or D.shape[1] != EMBED_DIM:
            raise EngineError(f"db_build: database must be [N,256], got {tuple(D.shape)}")
        self._db = D
        self._check(self._lib.t2l_db_build(self._h, _ptr(D), D.shape[0], int(row_offset), self._stream()))

    def search_topk(self, Q, k: int, exact: bool = False, out=None):
        """-> (idx int64 [nq,k], score float64 [nq,k], n_fallback int tensor [1]); order (score desc, row asc).
        out: optional pre-allocated (idx, score) pair to write into (e.g. the two halves of one packed buffer)."""
        if self._db is None:
            raise EngineError("search_topk: call db_build first")
        Q = self._dev(Q, torch.float32)
        if Q.dim() != 2 or Q.shape[1] != EMBED_DIM:
            raise EngineError(f"search_topk: queries must be [nq,256], got {tuple(Q.shape)}")
        nq = Q.shape[0]
        if out is not None:
            idx, sc = out
            if (idx.shape != (nq, k) or sc.shape != (nq, k) or idx.dtype != torch.int64 or sc.dtype != torch.float64
                    or not idx.is_contiguous() or not sc.is_contiguous()):
                raise EngineError("search_topk: out must be contiguous (int64 [nq,k], float64 [nq,k])")
        else:
            idx = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            sc = torch.empty((nq, k), dtype=torch.float64, device=self.device)
        nfb = torch.zeros(1, dtype=torch.int32, device=self.device)
        if exact or k > MAX_TOPK_FAST:
            self._check(self._lib.t2l_search_topk_exact(self._h, _ptr(Q), nq, k, _ptr(idx), _ptr(sc), self._stream()))
        else:
            self._check(self._lib.t2l_search_topk(self._h, _ptr(Q), nq, k, _ptr(idx), _ptr(sc), _ptr(nfb), self._stream()))
        return idx, sc, nfb

    def search_topk_accumulate(self, Q, k: int, run_idx: torch.Tensor, run_score: torch.Tensor):
        """Fold the top-k of the registered shard into the running lists (in place); returns n_fallback [1]."""
        if self._db is None:
            raise EngineError("search_topk_accumulate: call db_build first")
        Q = self._dev(Q, torch.float32)
        nq = Q.shape[0]
        if (run_idx.shape != (nq, k) or run_score.shape != (nq, k) or run_idx.dtype != torch.int64 or run_score.dtype != torch.float64
                or not run_idx.is_contiguous() or not run_score.is_contiguous() or run_idx.device != self.device):
            raise EngineError("search_topk_accumulate: running lists must be contiguous int64 / float64 [nq, k] on the engine's device")
        nfb = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self._lib.t2l_search_topk_accumulate(self._h, _ptr(Q), nq, k, _ptr(run_idx), _ptr(run_score), _ptr(nfb), self._stream()))
        return nfb

    def new_running_topk(self, nq: int, k: int):
        """Empty running lists for search_topk_accumulate: idx -1, score -inf."""
        return (torch.full((nq, k), -1, dtype=torch.int64, device=self.device),
                torch.full((nq, k), float("-inf"), dtype=torch.float64, device=self.device))

    def synth_cells(self, seed: int, first_cell: int, n_cells: int, obj_per_cell: int, out=None):
        """Counter-based synthetic cells on the device -> (pts [n,256,6], meta [n,7], cell_ptr np.int32 [n_cells+1])."""
        n = n_cells * obj_per_cell
        if out is None:
            out = (torch.empty((n, NUM_POINTS, 6), dtype=torch.float32, device=self.device),
                   torch.empty((n, 7), dtype=torch.float32, device=self.device))
        pts, meta = out[0][:n], out[1][:n]
        self._check(self._lib.t2l_synth_cells(self._h, int(seed), int(first_cell), n_cells, obj_per_cell, _ptr(pts), _ptr(meta), self._stream()))
        return pts, meta, (np.arange(n_cells + 1, dtype=np.int64) * obj_per_cell).astype(np.int32)

    def merge_topk(self, idx_all, score_all):
        """[G,nq,k] per-shard lists -> global (idx, score) [nq,k]."""
        idx_all = self._dev(idx_all, torch.int64)
        score_all = self._dev(score_all, torch.float64)
        G, nq, k = idx_all.shape
        idx = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        sc = torch.empty((nq, k), dtype=torch.float64, device=self.device)
        self._check(self._lib.t2l_merge_topk(self._h, _ptr(idx_all), _ptr(score_all), G, nq, k, _ptr(idx), _ptr(sc), self._stream()))
        return idx, sc

    def merge_topk_packed(self, packed_all: torch.Tensor, nq: int, k: int):
        """[G, 2, nq, k] int64 words (idx | score bits) gathered from G shards -> global (idx, score) [nq,k]."""
        G = packed_all.shape[0]
        if packed_all.dtype != torch.int64 or tuple(packed_all.shape[1:]) != (2, nq, k) or not packed_all.is_contiguous():
            raise EngineError("merge_topk_packed: need a contiguous int64 [G, 2, nq, k] buffer")
        idx = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        sc = torch.empty((nq, k), dtype=torch.float64, device=self.device)
        self._check(self._lib.t2l_merge_topk_packed(self._h, _ptr(packed_all), G, nq, k, _ptr(idx), _ptr(sc), self._stream()))
        return idx, sc

    # ---- bookkeeping ----------------------------------------------------------------------
    def topk_accuracy(self, idx, query_xy, cell_xy, top_k, threshs=(), target_row=None, query_scene=None, cell_scene=None,
                      want_dists: bool = False):
        """Per-query accuracy rows (t2l_topk_accuracy): returns (hit u8 [nq, n_top] or None, within u8 [nq, n_top, n_thr] or
        None, dists f64 [nq, k] or None), all on device."""
        idx = self._dev(idx, torch.int64)
        nq, k = idx.shape
        query_xy = self._dev(query_xy, torch.float64)
        cell_xy = self._dev(cell_xy, torch.float64)
        if query_xy.shape != (nq, 2) or cell_xy.dim() != 2 or cell_xy.shape[1] != 2:
            raise EngineError(f"topk_accuracy: query_xy {tuple(query_xy.shape)} / cell_xy {tuple(cell_xy.shape)}")
        tk = np.ascontiguousarray(np.asarray(list(top_k), dtype=np.int32))
        th = np.ascontiguousarray(np.asarray(list(threshs), dtype=np.float64))
        target_row = self._dev(target_row, torch.int64) if target_row is not None else None
        query_scene = self._dev(query_scene, torch.int32) if query_scene is not None else None
        cell_scene = self._dev(cell_scene, torch.int32) if cell_scene is not None else None
        hit = torch.empty((nq, len(tk)), dtype=torch.uint8, device=self.device) if target_row is not None else None
        within = torch.empty((nq, len(tk), len(th)), dtype=torch.uint8, device=self.device) if len(th) else None
        dists = torch.empty((nq, k), dtype=torch.float64, device=self.device) if want_dists else None
        with torch.cuda.device(self.device):
            self._check(self._lib.t2l_topk_accuracy(
                self._h, _ptr(idx), nq, k, _ptr(target_row), _ptr(query_xy), _ptr(cell_xy), _ptr(query_scene), _ptr(cell_scene),
                tk.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(tk), th.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(th),
                _ptr(hit), _ptr(within), _ptr(dists), self._stream()))
        return hit, within, dists

    # ---- test hooks -------------------------------------------------------------------------
    def debug_linear_f16(self, A, W, bias=None, act=0, out_half=False):
        """fp16-operand tcgen05 GEMM (the token layer's): A [M,K], W [N,K] float16 on device."""
        A, W = self._dev(A, torch.float16), self._dev(W, torch.float16)
        bias = self._dev(bias, torch.float32) if bias is not None else None
        M, K = A.shape
        N = W.shape[0]
        C = torch.empty((M, N), dtype=torch.float16 if out_half else torch.float32, device=self.device)
        self._check(self._lib.t2l_debug_linear_f16(self._h, _ptr(A), K, _ptr(W), K, _ptr(bias), _ptr(C), N, M, N, K, act,
                                                   int(out_half), self._stream()))
        return C

    def debug_linear_f16_residual(self, A, W, bias, R, reg_epilogue=False):
        """fp16(A W^T + bias + R): the out-projection / FFN2 of the token layer on its fp16 residual stream."""
        A, W, R = self._dev(A, torch.float16), self._dev(W, torch.float16), self._dev(R, torch.float16)
        bias = self._dev(bias, torch.float32) if bias is not None else None
        M, K = A.shape
        N = W.shape[0]
        buf = torch.full((M + 160, N), 7.0, dtype=torch.float16, device=self.device)  # the rows beyond M must stay untouched
        self._check(self._lib.t2l_debug_linear_f16_residual(self._h, _ptr(A), K, _ptr(W), K, _ptr(bias), _ptr(R), N, _ptr(buf), N, M, N, K,
                                                            int(reg_epilogue), self._stream()))
        if not bool((buf[M:] == 7.0).all()):
            raise RuntimeError("debug_linear_f16_residual: rows beyond M were written")
        return buf[:M]

    def debug_mha(self, qkv, n_seq: int, S: int, n_heads: int):
        """Small-sequence attention core on packed rows [n_seq*S, 3d] -> [n_seq*S, d]."""
        qkv = self._dev(qkv, torch.float32)
        d = qkv.shape[1] // 3
        out = torch.empty((n_seq * S, d), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_debug_mha(self._h, _ptr(qkv), _ptr(out), n_seq, S, d, n_heads, self._stream()))
        return out

    def debug_mha_cross(self, q, kv, n_seq: int, Sq: int, Sk: int, n_heads: int):
        """Cross attention core: q [n_seq*Sq, d], kv [n_seq*Sk, 2d] (k | v) -> [n_seq*Sq, d]."""
        q, kv = self._dev(q, torch.float32), self._dev(kv, torch.float32)
        d = q.shape[1]
        out = torch.empty((n_seq * Sq, d), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_debug_mha_cross(self._h, _ptr(q), _ptr(kv), _ptr(out), n_seq, Sq, Sk, d, n_heads, self._stream()))
        return out

    def debug_mha_cells(self, qkv, row_ptr, cell_ptr, slots: int, n_heads: int):
        """Intra-cell attention core on packed rows with one representative padding row per cell."""
        qkv = self._dev(qkv, torch.float32)
        row_ptr, cell_ptr = self._dev(row_ptr, torch.int32), self._dev(cell_ptr, torch.int32)
        d = qkv.shape[1] // 3
        out = torch.empty((qkv.shape[0], d), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_debug_mha_cells(self._h, _ptr(qkv), _ptr(out), row_ptr.numel() - 1, _ptr(row_ptr), _ptr(cell_ptr), slots, d,
                                                  n_heads, self._stream()))
        return out

    def debug_linear(self, A, W, bias=None, act=0, segmax=False, path=1):
        rowmajor = lambda t: t if (t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1) else self._dev(t, torch.float32)
        A, W = rowmajor(A), rowmajor(W)  # row-padded views (stride(0) > K) are passed through as lda / ldw
        bias = self._dev(bias, torch.float32) if bias is not None else None
        M, K = A.shape
        N = W.shape[0]
        C = torch.empty((M // 32 if segmax else M, N), dtype=torch.float32, device=self.device)
        self._check(self._lib.t2l_debug_linear(self._h, path, _ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(bias), _ptr(C), N,
                                               M, N, K, act, int(segmax), self._stream()))
        return C

"""World-size-2 gloo test of the multi-GPU host logic (row sharding, all-gather, merge order) with
the numpy oracle standing in for the per-shard device search."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import restate
    from text2loc_b200 import distributed as t2ld
    import synth

    D = synth.make_unit_rows(3, 1001)  # not divisible by the world size
    D[900] = D[5]  # an exact tie across the two shards
    Q = np.concatenate([D[5:6], synth.make_unit_rows(4, 37)])
    lo, hi = t2ld.shard_bounds(len(D), world, rank)
    idx, sc = restate.search_topk(D[lo:hi], Q, 10)  # what t2l_search_topk returns for this shard
    idx_all = t2ld.all_gather_rows(torch.from_numpy(idx + lo))
    sc_all = t2ld.all_gather_rows(torch.from_numpy(sc))
    midx, msc = t2ld.merge_topk_host(idx_all.numpy(), sc_all.numpy(), 10)
    oidx, osc = restate.search_topk(D, Q, 10)
    ok = bool((midx == oidx).all() and np.array_equal(msc, osc) and midx[0, 0] == 5 and midx[0, 1] == 900)
    # query embeddings split by rank and gathered back in rank order
    mine = torch.full((3, 4), float(rank))
    g = t2ld.all_gather_rows(mine).reshape(-1, 4)
    ok = ok and g.shape == (3 * world, 4) and bool((g[:3] == 0).all() and (g[3:6] == 1).all())
    # the production exchange: ONE all-gather of the packed (idx | score bits) buffer, merged in place -- sharded_search with the
    # numpy oracle behind the engine's two calls
    class OracleEngine:
        def search_topk(self, Qt, k, out=None):
            i, s_ = restate.search_topk(D[lo:hi], Qt.numpy(), k)
            out[0].copy_(torch.from_numpy(i + lo))
            out[1].copy_(torch.from_numpy(s_))
            return out[0], out[1], torch.zeros(1, dtype=torch.int32)

        def merge_topk_packed(self, gathered, nq, k):
            assert gathered.shape == (world, 2, nq, k) and gathered.dtype == torch.int64
            i, s_ = t2ld.merge_topk_host(gathered[:, 0].numpy(), gathered[:, 1].contiguous().view(torch.float64).numpy(), k)
            return torch.from_numpy(i), torch.from_numpy(s_)

    Qs = torch.from_numpy(Q)
    per = (len(Q) + world - 1) // world  # queries split by rank (padded to equal shares), gathered inside sharded_search
    Qpad = torch.cat([Qs, Qs[:per * world - len(Q)]])
    pidx, psc, _ = t2ld.sharded_search(OracleEngine(), Qpad[rank * per:(rank + 1) * per], 10, queries_are_sharded=True)
    ok = ok and bool((pidx[:len(Q)].numpy() == oidx).all() and np.array_equal(psc[:len(Q)].numpy(), osc))
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def test_sharded_merge_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    bounds = dict((r, b) for r, _, b in res)
    assert bounds[0] == (0, 501) and bounds[1] == (501, 1001)


def test_shard_bounds_cover_rows_exactly():
    from text2loc_b200.distributed import shard_bounds

    for n in (0, 1, 7, 100000, 12345):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_merge_host_handles_empty_slots():
    from text2loc_b200.distributed import merge_topk_host

    idx = np.array([[[4, 2, -1]], [[7, -1, -1]]])
    sc = np.array([[[0.9, 0.5, -np.inf]], [[0.9, -np.inf, -np.inf]]])
    i, s = merge_topk_host(idx, sc, 3)
    assert i.tolist() == [[4, 7, 2]] and s.tolist() == [[0.9, 0.9, 0.5]]

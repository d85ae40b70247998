"""N-rank NCCL check: row-sharded search + all-gather + merge equals the single-GPU search and the fp64 oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text2loc_b200 import distributed as t2ld  # noqa: E402
import synth  # noqa: E402
from text2loc_b200.engine import Engine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(f"cuda:{local}")
D = synth.make_unit_rows(31, 50001)
D[40000] = D[3]  # exact tie across shards
Q = np.concatenate([D[3:4], synth.make_unit_rows(32, 1023)])
lo, hi = t2ld.shard_bounds(len(D), world, rank)
eng.db_build(D[lo:hi], row_offset=lo)
# queries sharded by rank (as after a query-split text head), gathered inside sharded_search
qlo, qhi = t2ld.shard_bounds(len(Q), world, rank)
assert (qhi - qlo) * world == len(Q)
idx, sc, nfb = t2ld.sharded_search(eng, torch.from_numpy(Q[qlo:qhi]).cuda(), 10, queries_are_sharded=True)
torch.cuda.synchronize()
if rank == 0:
    from oracle import restate

    oidx, osc = restate.search_topk(D, Q, 10)
    ok = (idx.cpu().numpy() == oidx).all() and np.abs(sc.cpu().numpy() - osc).max() < 1e-12
    one = Engine(f"cuda:{local}")
    one.db_build(D)
    i1, s1, _ = one.search_topk(Q, 10)
    ok = ok and torch.equal(i1, idx) and torch.equal(s1, sc)
    print(f"world {world}: sharded == single-GPU == fp64 oracle: {bool(ok)}; first row {idx[0, :3].tolist()}")
    assert ok
dist.barrier()
dist.destroy_process_group()

"""Search-only profiling driver (32768 queries x N rows)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth  # noqa: E402
from text2loc_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
eng = Engine("cuda:0")
D = torch.from_numpy(synth.make_unit_rows(77, n)).cuda()
Q = torch.from_numpy(synth.make_unit_rows(78, 32768)).cuda()
eng.db_build(D)
eng.search_topk(Q, 10)
torch.cuda.synchronize()
torch.cuda.profiler.start()
idx, sc, nfb = eng.search_topk(Q, 10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("fallbacks", int(nfb))

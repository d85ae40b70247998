"""Search on ENCODER embeddings (random-weight encoders give near-collinear rows: the hard case for the proof)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth  # noqa: E402
from text2loc_b200.engine import Engine  # noqa: E402

eng = Engine("cuda:0")
eng.load_state_dict(synth.make_state_dict(0))
pts, meta, ptr = synth.make_packed_cells(1001, 12500, 8)
D = eng.encode_cells(torch.from_numpy(pts).cuda(), torch.from_numpy(meta).cuda(), ptr)
q = eng.encode_text(torch.from_numpy(synth.make_t5_features(2001, 4096)).cuda(), 6)
Q = q.repeat(8, 1).contiguous()
S = (Q[:64].double() @ D.double().T)
top = S.topk(17, dim=1).values
print("score range", float(S.min()), float(S.max()), "median gap rank10-rank16", float((top[:, 9] - top[:, 15]).median()))
eng.db_build(D)
eng.search_topk(Q, 10)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
idx, sc, nfb = eng.search_topk(Q, 10)
b.record()
torch.cuda.synchronize()
print("search 32768 x 12500 encoder embeddings:", a.elapsed_time(b), "ms; fallbacks", int(nfb))
torch.cuda.profiler.start()
eng.search_topk(Q, 10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

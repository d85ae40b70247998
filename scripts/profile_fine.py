"""Profiling driver for the fine stage (CrossMatch): one batched pass inside a cudaProfilerStart/Stop range."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

nq, top, n_cells, pad = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 5, 2000, 16
eng = Engine("cuda:0")
eng.load_state_dict(synth.make_fine_state_dict(0))
pts, meta, ptr = eng.synth_cells(3, 0, n_cells, pad)
t5 = torch.from_numpy(synth.make_t5_features(4, nq, 6, 12)).cuda()
rng = np.random.default_rng(0)
pair_cell = torch.from_numpy(rng.integers(0, n_cells, nq * top).astype(np.int32)).cuda()
pair_query = torch.arange(nq, dtype=torch.int32, device="cuda").repeat_interleave(top)
def run():
    obj = eng.fine_encode_objects(pts, meta, ptr)
    hints = eng.fine_encode_hints(t5)
    return eng.fine_match(obj, pair_cell, hints, pair_query, pad, 6)
run(); torch.cuda.synchronize()
obj = eng.fine_encode_objects(pts, meta, ptr); hints = eng.fine_encode_hints(t5); torch.cuda.synchronize()
torch.cuda.profiler.start()
off = eng.fine_match(obj, pair_cell, hints, pair_query, pad, 6)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("pairs", nq * top, "finite", bool(torch.isfinite(off).all()))

"""Encode throughput of BASELINE configs[1]'s database (10 000 cells x 8 objects) for one T2L_OBJ_CHUNK (env)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

eng = Engine("cuda:0")
eng.load_state_dict(synth.make_state_dict(0))
pts, meta, ptr = synth.make_packed_cells(1000, 10000, 8)
pts, meta = torch.from_numpy(pts).cuda(), torch.from_numpy(meta).cuda()
ms = []
for _ in range(5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.encode_cells(pts, meta, ptr)
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
print(f"chunk {os.environ.get('T2L_OBJ_CHUNK', 'default')}: runs {[round(m, 1) for m in ms]} ms -> {10000 / np.median(ms[1:]) * 1e3:.0f} cells/s")

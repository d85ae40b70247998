"""Search-only timing (32768 queries x N rows), CUDA events, both first-pass modes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

for mode in ("fp16", "bf16x3"):
    os.environ["T2L_SEARCH_FIRST"] = mode
    eng = Engine("cuda:0")
    for n in (12500, 100000):
        D = torch.from_numpy(synth.make_unit_rows(77, n)).cuda()
        Q = torch.from_numpy(synth.make_unit_rows(78, 32768)).cuda()
        eng.db_build(D)
        for _ in range(3):
            eng.search_topk(Q, 10)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            idx, sc, nfb = eng.search_topk(Q, 10)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"first pass {mode}: 32768 x {n}: {ms:.3f} ms, {2*32768*n*256/ms/1e9:.0f} TFLOP/s algorithmic, second-pass queries {int(nfb)}")

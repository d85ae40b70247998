"""Timing bisect of sa_obj2 (profiling only): which of accumulator drain / MMAs / gather bounds each level.
   ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:sa_obj2 ... python scripts/sa_bisect.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

eng = Engine("cuda:0")
eng.load_state_dict(synth.make_state_dict(0))
pts, meta, ptr = synth.make_packed_cells(1, 512, 8)
pts, meta = torch.from_numpy(pts).cuda(), torch.from_numpy(meta).cuda()
eng.encode_cells(pts, meta, ptr)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for mode in (0, 1, 2, 3):
    eng._check(eng._lib.t2l_debug_sa_bisect(eng._h, mode))
    eng.encode_cells(pts, meta, ptr)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
eng._check(eng._lib.t2l_debug_sa_bisect(eng._h, 0))

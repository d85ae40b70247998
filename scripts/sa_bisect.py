"""Timing bisect of sa_fused_kernel: run one encode chunk per T2L_SA_DBG mask inside the profiler range (results are wrong for mask != 0)."""
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
for m in [int(x) for x in sys.argv[1:]]:
    os.environ["T2L_SA_DBG"] = str(m)
    eng.encode_cells(pts, meta, ptr)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()

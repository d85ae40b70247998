"""Profiling driver: one DB-encode chunk + one query step inside a cudaProfilerStart/Stop range.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -o gpurun_out/prof python scripts/profile_step.py --small
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth  # noqa: E402
from text2loc_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true", help="one chunk of each stage only (for --set full captures)")
ap.add_argument("--skip-cells", action="store_true", help="profile the query step only")
ap.add_argument("--cells", type=int, default=0)
ap.add_argument("--queries", type=int, default=0)
args = ap.parse_args()
n_cells = args.cells or (256 if args.small else 10000)
nq = args.queries or (455 if args.small else 4096)

eng = Engine("cuda:0")
eng.load_state_dict(synth.make_state_dict(0))
pts, meta, ptr = synth.make_packed_cells(1, n_cells, 8)
pts, meta = torch.from_numpy(pts).cuda(), torch.from_numpy(meta).cuda()
t5 = torch.from_numpy(synth.make_t5_features(2, nq)).cuda()
D = eng.encode_cells(pts, meta, ptr)  # warm-up (allocates the arena)
eng.db_build(D)
q = eng.encode_text(t5, 6)
eng.search_topk(q, 10)
torch.cuda.synchronize()

torch.cuda.profiler.start()
if not args.skip_cells:
    D = eng.encode_cells(pts, meta, ptr)
    eng.db_build(D)
q = eng.encode_text(t5, 6)
idx, sc, nfb = eng.search_topk(q, 10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", n_cells, "cells", nq, "queries; fallbacks", int(nfb))

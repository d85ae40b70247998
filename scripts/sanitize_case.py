"""Small-shape pass over every kernel of the path for compute-sanitizer (scripts/gpu_sanitize.sh)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

which = sys.argv[1] if len(sys.argv) > 1 else "all"
sd = synth.make_state_dict(0)
eng = Engine("cuda:0"); eng.load_state_dict(sd)
if which in ("all", "cells"):
    pts, meta, ptr = synth.make_packed_cells(41, 40, 8)  # 320 objects: two to three objects per CTA of the fused SA kernels
    D = eng.encode_cells(pts, meta, ptr)
    torch.cuda.synchronize()
    print("cells ok", float(D.norm(dim=1).mean()))
if which in ("all", "text"):
    q = eng.encode_text(torch.from_numpy(synth.make_t5_features(3, 16)).cuda(), 6)
    torch.cuda.synchronize()
    print("text ok", float(q.norm(dim=1).mean()))
if which in ("all", "search"):
    Dn = torch.from_numpy(synth.make_unit_rows(1, 3000)).cuda()
    Qn = torch.from_numpy(synth.make_unit_rows(2, 300)).cuda()
    eng.db_build(Dn)
    idx, sc, nfb = eng.search_topk(Qn, 10)
    run = eng.new_running_topk(300, 10)
    eng.search_topk_accumulate(Qn, 10, *run)
    torch.cuda.synchronize()
    print("search ok", int(nfb), bool((run[0] == idx).all()))
if which in ("all", "fine"):
    # fine stage (CrossMatch): object encoder at d = 128, hint encoder, decoder layers (sequence-per-warp self / cross attention,
    # operand planes written by LayerNorm / attention / the ReLU epilogue), offset MLP
    fe = Engine("cuda:0"); fe.load_state_dict(synth.make_fine_state_dict(0))
    pts, meta, ptr = fe.synth_cells(3, 0, 12, 16)
    t5 = torch.from_numpy(synth.make_t5_features(4, 12, 6, 12)).cuda()
    off = fe.fine_offsets(pts, meta, ptr, t5, 6)
    torch.cuda.synchronize()
    print("fine ok", bool(torch.isfinite(off).all()))

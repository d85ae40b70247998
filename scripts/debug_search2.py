"""Debug driver for the second search pass: prints which queries differ from the fp64 oracle."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restate
import synth
from text2loc_b200.engine import Engine

eng = Engine("cuda:0")
rng = np.random.default_rng(3)
base = synth.make_unit_rows(40, 1)[0]
D = (base[None, :] + 0.02 * rng.standard_normal((6000, 256))).astype(np.float32)
D /= np.linalg.norm(D, axis=1, keepdims=True)
D[1000:1300] = D[999]
Q = np.concatenate([D[999:1000], (base[None, :] + 0.02 * rng.standard_normal((200, 256))).astype(np.float32)])
for name, DD, QQ in (("dups", D, Q), ("nodups", np.concatenate([D[:1000], D[1300:]]), Q[1:]), ("dups_1q", D, Q[:1]), ("dups_q107", D, Q[100:120])):
    eng.db_build(DD)
    idx, sc, nfb = eng.search_topk(QQ, 10)
    eidx, esc, _ = eng.search_topk(QQ, 10, exact=True)
    oidx, osc = restate.search_topk(DD, QQ, 10)
    idx, eidx = idx.cpu().numpy(), eidx.cpu().numpy()
    bad = np.where((idx != oidx).any(axis=1))[0]
    bad_e = np.where((eidx != oidx).any(axis=1))[0]
    print(f"[{name}] nq={len(QQ)} N={len(DD)} nfb={int(nfb)} mismatching queries fast={bad.tolist()} exact={bad_e.tolist()}")
    for q in bad[:4]:
        print("  q", q, "engine", idx[q].tolist(), "\n        oracle", oidx[q].tolist(), "\n        exact ", eidx[q].tolist())
        print("     eng score", sc[q].cpu().numpy().tolist()[:4], "oracle", osc[q].tolist()[:4])

"""Which objects does sa_obj2 get wrong? features2 of the default engine vs the T2L_SA_V1 engine, per object."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

sd = synth.make_state_dict(0)
for n_cells, per in ((20, 8), (37, 8), (40, 8), (150, 8), (300, 16)):
    pts, meta, ptr = synth.make_packed_cells(41, n_cells, per)
    os.environ.pop("T2L_SA_V1", None)
    e2 = Engine("cuda:0"); e2.load_state_dict(sd)
    a = e2.encode_objects_debug(torch.from_numpy(pts), ptr)["features2"].cpu().numpy()
    os.environ["T2L_SA_V1"] = "1"
    e1 = Engine("cuda:0"); e1.load_state_dict(sd)
    b = e1.encode_objects_debug(torch.from_numpy(pts), ptr)["features2"].cpu().numpy()
    os.environ.pop("T2L_SA_V1", None)
    err = np.abs(a - b).max(axis=1) / np.abs(b).max()
    bad = np.nonzero(err > 2e-3)[0]
    n = n_cells * per
    print(f"n_obj {n}: max err {err.max():.2e}, bad objects {len(bad)}: {bad[:40].tolist()}")
    if len(bad):
        # position inside the CTA's run of objects
        grid = min(n, 148)
        runs = [(int(n * c / grid), int(n * (c + 1) / grid)) for c in range(grid)]
        pos = []
        for o in bad[:40]:
            for (s, t) in runs:
                if s <= o < t:
                    pos.append((o - s, t - s)); break
        print("   (index in its CTA's run, run length):", pos)

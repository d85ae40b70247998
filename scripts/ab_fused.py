"""A/B: fused set-abstraction kernel vs the v1 gather -> H -> GEMM path (bitwise) + encode timing."""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth  # noqa: E402
from text2loc_b200.engine import Engine  # noqa: E402

if len(sys.argv) > 1:  # child: unfused features
    eng = Engine("cuda:0")
    eng.load_state_dict(synth.make_state_dict(0))
    pts, meta, ptr = synth.make_packed_cells(5, 64, 8)
    np.save(sys.argv[1], eng.encode_objects_debug(pts, ptr)["features2"].cpu().numpy())
    sys.exit(0)

eng = Engine("cuda:0")
eng.load_state_dict(synth.make_state_dict(0))
pts, meta, ptr = synth.make_packed_cells(5, 64, 8)
f_fused = eng.encode_objects_debug(pts, ptr)["features2"].cpu().numpy()
subprocess.run([sys.executable, __file__, "/tmp/unfused.npy"], check=True, env={**os.environ, "T2L_UNFUSED_SA": "1"})
f_un = np.load("/tmp/unfused.npy")
print("fused vs unfused features2: max abs diff", np.abs(f_fused - f_un).max(), "bitwise equal:", np.array_equal(f_fused, f_un))
pts, meta, ptr = synth.make_packed_cells(6, 2560, 8)
pts, meta = torch.from_numpy(pts).cuda(), torch.from_numpy(meta).cuda()
for _ in range(2):
    eng.encode_cells(pts, meta, ptr)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    eng.encode_cells(pts, meta, ptr)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"encode 2560 cells x 8 objects: {ms:.2f} ms -> {2560 / ms * 1e3:.0f} cells/s, {2560 * 8 * 377.7e6 / (ms * 1e-3) / 1e12:.1f} TFLOP/s algorithmic")

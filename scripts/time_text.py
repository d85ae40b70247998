"""Text head timing: fp32 vs fp16 T5 features (device-resident), interleaved."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

eng = Engine("cuda:0"); eng.load_state_dict(synth.make_state_dict(0))
t5 = torch.from_numpy(synth.make_t5_features(2, 4096)).cuda()
t16 = t5.half()
res = {"f32": [], "f16": []}
for _ in range(3):
    eng.encode_text(t5, 6); eng.encode_text(t16, 6)
for rep in range(6):
    for name, x in (("f32", t5), ("f16", t16)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            eng.encode_text(x, 6)
        b.record(); torch.cuda.synchronize()
        res[name].append(a.elapsed_time(b) / 5)
print({k: [round(v, 3) for v in vs] for k, vs in res.items()})
if len(sys.argv) > 1:
    torch.cuda.profiler.start()
    eng.encode_text(t16 if sys.argv[1] == "f16" else t5, 6)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()

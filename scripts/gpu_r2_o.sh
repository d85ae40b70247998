#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search or shard or merge or topk or stream or text or e2e or host" 2>&1 | tail -n 3
timeout 300 python scripts/time_search.py 2>&1 | head -n 2
timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import synth
from text2loc_b200.engine import Engine
eng = Engine("cuda:0"); eng.load_state_dict(synth.make_state_dict(0))
t5 = torch.from_numpy(synth.make_t5_features(2, 4096)).half().pin_memory()
dev = torch.empty_like(t5, device="cuda")
def tm(fn, n=5):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b), 3))
    return ts
print("pure H2D of the 604 MB, one copy:", tm(lambda: dev.copy_(t5, non_blocking=True)))
ref = eng.encode_text(dev, 6)
for _ in range(3): eng.encode_text(t5, 6)
print("e2e text (host fp16 -> embeddings) ms:", tm(lambda: eng.encode_text(t5, 6)))
got = eng.encode_text(t5, 6)
print("host-streamed == device-resident:", bool(torch.equal(got, ref)), float((got - ref).abs().max()))
PY

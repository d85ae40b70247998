#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search or shard or merge or topk or stream" 2>&1 | tail -n 4
timeout 300 python scripts/time_search.py 2>&1 | tail -n 6
for c in 16384 32768 65536; do
T2L_HOST_TOK_CHUNK=$c timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import synth
from text2loc_b200.engine import Engine
eng = Engine("cuda:0"); eng.load_state_dict(synth.make_state_dict(0))
t5 = torch.from_numpy(synth.make_t5_features(2, 4096)).half().pin_memory()
for _ in range(3): eng.encode_text(t5, 6)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); q = eng.encode_text(t5, 6); b.record(); torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b), 3))
print("host chunk", os.environ["T2L_HOST_TOK_CHUNK"], "e2e text ms", ts)
PY
done

#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --skip-extras > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?"
[ "${STREAM:-1}" = "1" ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload stream > $O/bench_stream_n$N.json 2> $O/bench_stream_n$N.err; echo "stream n$N rc=$?"
tail -c 300 $O/bench_n$N.err; tail -c 300 $O/bench_stream_n$N.err
python - <<PY
import json
for f in ("bench_n$N.json", "bench_stream_n$N.json"):
    ls=[l for l in open("$O/"+f) if l.startswith("{")]
    if not ls: print(f, "no json"); continue
    d=json.loads(ls[-1])
    print(f, {k:d.get(k) for k in ("value","ms_per_step","ms_text_head","ms_search","stage_ms","cells_per_s","topk_matches_fp64_oracle_sample","search_fallbacks")}, "e2e", (d.get("e2e") or {}).get("value"), d.get("clocks"))
PY

#!/bin/bash
# multi-GPU pass: N ranks over NCCL; checks the sharded result against a 1-GPU search on rank 0
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/check_sharded.py > gpurun_out/sharded_$N.log 2>&1; echo "sharded check rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -n 5 gpurun_out/sharded_$N.log; tail -c 1500 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.log

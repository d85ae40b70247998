#!/bin/bash
# first GPU pass: kernel tests, parity tests, smoke, short bench -- each in its own process
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
for t in test_umma_gemm_matches_tf32_emulation test_umma_segmax_epilogue test_simt_gemm_fp32 test_fps_and_ball_query_bit_exact test_features2_within_tolerance; do
  timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k $t > gpurun_out/k_$t.log 2>&1; echo "kernels/$t rc=$?"
done
for t in test_encode_cells test_encode_text test_search test_merge test_dropin; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k $t > gpurun_out/p_$t.log 2>&1; echo "parity/$t rc=$?"
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/*.log | tail -n 120

#!/bin/bash
# 2-GPU session: bench at N=2 (weak + strong sub-record + single-process multi-GPU), both arms; multi-GPU pytest
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.log 2>&1
echo "rc=$?" >> gpurun_out/r2_bench_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_benchref_n2.log 2>&1
echo "rc=$?" >> gpurun_out/r2_benchref_n2.log
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "multi_gpu or staged or resume" > gpurun_out/r2_pytest9.log 2>&1
tail -c 1500 gpurun_out/r2_bench_n2.log; echo; tail -c 600 gpurun_out/r2_benchref_n2.log; tail -3 gpurun_out/r2_pytest9.log

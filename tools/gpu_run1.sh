#!/bin/bash
# round-2 GPU session 1: full GPU test suite, then short bench lines for c1 and c4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c1_a.log 2>&1
timeout 400 python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/r2_bench_c4_a.log 2>&1
tail -3 gpurun_out/r2_pytest1.log
tail -c 600 gpurun_out/r2_bench_c1_a.log
tail -c 600 gpurun_out/r2_bench_c4_a.log

#!/bin/bash
# RTIOW 1080p (BASELINE config 2) on 8 GPUs with the last build: weak scaling, the fixed image as `strong` sub-record
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR bench.py --workload c2 --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/r2b_c2_n8.json 2> $O/r2b_c2_n8.err
grep "^{" $O/r2b_c2_n8.json | head -c 300; echo; tail -n 1 $O/r2b_c2_n8.err

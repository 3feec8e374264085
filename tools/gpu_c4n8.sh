#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --workload c4 --scaling strong --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_final_c4_strong_n$N.json 2> $O/r2_final_c4_strong_n$N.err
grep "^{" $O/r2_final_c4_strong_n$N.json | head -c 400; tail -n 2 $O/r2_final_c4_strong_n$N.err

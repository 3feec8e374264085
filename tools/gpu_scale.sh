#!/bin/bash
# scaling points between 1 and 8 GPUs: default scene (weak, with the fixed image as `strong` sub-record) and config 5 as a fixed image
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_final_c1_n$N.json 2> $O/r2_final_c1_n$N.err
timeout 900 $TR bench.py --workload c5 --scaling strong --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_final_c5_strong_n$N.json 2> $O/r2_final_c5_strong_n$N.err
for f in c1_n$N c5_strong_n$N; do echo "== $f"; grep "^{" $O/r2_final_$f.json | head -c 300; echo; tail -n 1 $O/r2_final_$f.err; done

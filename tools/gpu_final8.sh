#!/bin/bash
# round-2 final evidence on 8 GPUs: the default scene (weak scaling + the fixed image as `strong` sub-record), the
# reference arm under torchrun, and BASELINE config 5 -- the north-star's 8-GPU configuration -- as a fixed image
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_final_c1_n$N.json 2> $O/r2_final_c1_n$N.err
timeout 900 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/r2_final_ref_c1_n$N.json 2> $O/r2_final_ref_c1_n$N.err
timeout 900 $TR bench.py --workload c5 --scaling strong --gpus $N --steps 2 --warmup 3 > $O/r2_final_c5_strong_n$N.json 2> $O/r2_final_c5_strong_n$N.err
timeout 900 $TR bench.py --workload c4 --scaling strong --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_final_c4_strong_n$N.json 2> $O/r2_final_c4_strong_n$N.err
for f in c1_n$N ref_c1_n$N c5_strong_n$N c4_strong_n$N; do echo "== $f"; head -c 400 $O/r2_final_$f.json; echo; tail -n 2 $O/r2_final_$f.err; done

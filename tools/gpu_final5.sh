#!/bin/bash
# the last build on 2 GPUs: the whole GPU suite (with the single-process 2-GPU test), c1 and c2 under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 python -m pytest tests -m gpu -x -q -rs > $O/r2b_pytest_gpu_n2.log 2>&1
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r2b_c1_n2.json 2> $O/r2b_c1_n2.err
timeout 600 $TR bench.py --workload c2 --gpus $N --steps 10 --warmup 3 > $O/r2b_c2_n2.json 2> $O/r2b_c2_n2.err
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/r2b_ref_c1_n2.json 2> $O/r2b_ref_c1_n2.err
tail -4 $O/r2b_pytest_gpu_n2.log
for f in c1_n2 c2_n2 ref_c1_n2; do echo "== $f"; grep "^{" $O/r2b_$f.json | head -c 300; echo; tail -n 1 $O/r2b_$f.err; done

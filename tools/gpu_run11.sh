#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run11.log; : > $L
for v in r1 cur noinorder t768 t832 head4 cur; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
timeout 600 python -m pytest tests/test_rays_gpu.py -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1
tail -3 gpurun_out/r2_pytest11.log >> $L
for c in c2 c3; do timeout 200 python tools/express_sweep.py $c 0 -1 >> $L 2>&1; done
timeout 200 python tools/express_sweep.py c5 32 -1 >> $L 2>&1
timeout 200 python tools/express_sweep.py c4 8 -1 >> $L 2>&1
cat $L

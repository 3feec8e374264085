#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run10.log; : > $L
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_pytest10.log 2>&1
tail -12 gpurun_out/r2_pytest10.log >> $L
for v in build/variants/r1.so path_tracer_b200/lib/libptb200.so; do timeout 120 python tools/variant_time.py $v c1 100 5 >> $L 2>&1; done
timeout 300 python tools/express_sweep.py c1 0 -1 6 10 14 17 20 24 >> $L 2>&1
timeout 300 python tools/express_sweep.py c2 0 -1 2 4 8 12 17 >> $L 2>&1
timeout 300 python tools/express_sweep.py c3 64 -1 2 4 8 17 28 40 >> $L 2>&1
timeout 300 python tools/express_sweep.py c5 32 -1 2 4 8 17 >> $L 2>&1
timeout 300 python tools/express_sweep.py c4 8 -1 2 4 8 17 >> $L 2>&1
cat $L

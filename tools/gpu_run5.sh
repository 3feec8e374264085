#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run5.log; : > $L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1
tail -3 gpurun_out/r2_pytest5.log >> $L
timeout 300 python tools/express_sweep.py c1 0 -1 4 8 12 17 24 32 >> $L 2>&1
timeout 300 python tools/express_sweep.py c2 0 -1 8 17 >> $L 2>&1
cat $L

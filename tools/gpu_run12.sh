#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run12.log; : > $L
for v in r1 cur2 head4 cur2; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest12.log 2>&1
tail -5 gpurun_out/r2_pytest12.log >> $L
cat $L

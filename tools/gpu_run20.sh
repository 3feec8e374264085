#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=gpurun_out/r2_run20.log; : > $L
for v in s1 g1 s1 r1; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
for v in s1 g1; do timeout 120 python tools/variant_time.py build/variants/$v.so c2 64 3 >> $L 2>&1; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest20.log 2>&1; tail -3 gpurun_out/r2_pytest20.log >> $L
cat $L

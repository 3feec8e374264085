#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=gpurun_out/r2_run13.log; : > $L
for v in cur2 x0 x1 x2 x12 head4 cur2; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
cat $L

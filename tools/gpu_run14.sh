#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=gpurun_out/r2_run14.log; : > $L
for v in $VARIANTS; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
cat $L

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=gpurun_out/r2_run16.log; : > $L
for v in $VARIANTS; do timeout 200 python tools/variant_time.py build/variants/$v.so c4 16 3 >> $L 2>&1; done
cat $L

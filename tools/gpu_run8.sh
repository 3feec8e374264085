#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run8.log; : > $L
PT_PHASE_TIMING=1 timeout 120 python tools/phase_compare.py build/variants/phaseall.so c1 >> $L 2>&1
PT_PHASE_TIMING=1 timeout 120 python tools/phase_compare.py build/variants/phaseall.so c3 64 >> $L 2>&1
for v in r1 head4 r1 head4; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
for v in r1 head4; do timeout 120 python tools/variant_time.py build/variants/$v.so c3 64 3 >> $L 2>&1; done
cat $L

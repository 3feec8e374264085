#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run4.log; : > $L
for v in r1 head2 r1 head2; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
for v in r1 head2; do timeout 120 python tools/variant_time.py build/variants/$v.so c3 64 3 >> $L 2>&1; done
echo "--- express-only phase timing" >> $L
PT_PHASE_TIMING=1 PTB200_LIB=build/variants/phase.so timeout 120 python tools/timeline.py 100 >> $L 2>&1
echo "--- all-CTA phase timing" >> $L
PT_PHASE_TIMING=1 PTB200_LIB=build/variants/phaseall.so timeout 120 python tools/timeline.py 100 >> $L 2>&1
cat $L

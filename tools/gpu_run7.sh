#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run7.log; : > $L
for v in r1phase phaseall; do PT_PHASE_TIMING=1 timeout 120 python tools/phase_compare.py build/variants/$v.so c1 >> $L 2>&1; done
for v in r1phase phaseall; do PT_PHASE_TIMING=1 timeout 120 python tools/phase_compare.py build/variants/$v.so c3 64 >> $L 2>&1; done
for v in r1 head3; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
cat $L

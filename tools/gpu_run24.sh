#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PT_PHASE_TIMING=1 python tools/phase_compare.py build/variants/pht.so c4 16
PT_PHASE_TIMING=1 python tools/phase_compare.py build/variants/pht.so c3 64
PT_PHASE_TIMING=1 python tools/phase_compare.py build/variants/pht.so c2 64
} > $O/r2_run24.log 2>&1
cat $O/r2_run24.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PT_LAST_PIXEL=1 PTB200_LIB=build/variants/lastpx.so python tools/express_sweep.py c4 16 -1
} > $O/r2_run38.log 2>&1
head -50 $O/r2_run38.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PT_LAST_PIXEL=1 PTB200_LIB=build/variants/lastpx.so python tools/tail_probe.py c4 64
PT_LAST_PIXEL=1 PTB200_LIB=build/variants/lastpx.so python tools/tail_probe.py c3 64
} > $O/r2_run34.log 2>&1
cat $O/r2_run34.log

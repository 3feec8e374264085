#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PTB200_LIB=build/variants/x2.so python tools/express_sweep.py c4 64 -1
PTB200_LIB=build/variants/x2.so python tools/express_sweep.py c3 64 -1
} > $O/r2_run37.log 2>&1
cat $O/r2_run37.log

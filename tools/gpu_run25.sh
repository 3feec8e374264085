#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
python tools/express_sweep.py c4 64 -1 8 24
python tools/express_sweep.py c4 256 -1
} > $O/r2_run25.log 2>&1
cat $O/r2_run25.log

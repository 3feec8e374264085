#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
echo "--- hand-off off"; PTB200_LIB=build/variants/hrinf.so python tools/tail_probe.py c4 64
echo "--- hand-off on (q16)"; PTB200_LIB=build/variants/q16.so python tools/tail_probe.py c4 64
} > $O/r2_run27.log 2>&1
cat $O/r2_run27.log

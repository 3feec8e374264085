#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PTB200_LIB=build/variants/ps1.so python tools/tail_probe.py c4 64
for c in "c1 100" "c2 64" "c3 64" "c5 16"; do python tools/variant_time.py build/variants/ps1.so $c 3; done
} > $O/r2_run31.log 2>&1
cat $O/r2_run31.log

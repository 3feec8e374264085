#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PTB200_LIB=build/variants/adapt.so python tools/tail_probe.py c4 64
for v in l64 l96; do python tools/variant_time.py build/variants/$v.so c4 64 3; done
} > $O/r2_run30.log 2>&1
cat $O/r2_run30.log

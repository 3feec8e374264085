#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in t00 t10 t01 t11; do python tools/variant_time.py build/variants/$v.so c4 16 3; done
for v in t00 t11; do python tools/variant_time.py build/variants/$v.so c3 64 3; done
PTB200_LIB=build/variants/t11.so timeout 900 python -m pytest tests -m gpu -x -q -k "flat or config4 or mesh or tree" 2>&1 | tail -3
} > $O/r2_run22.log 2>&1
cat $O/r2_run22.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in q100000 q64 q16 q4; do
  for c in "c1 100" "c2 64" "c3 64" "c4 64" "c5 16"; do python tools/variant_time.py build/variants/$v.so $c 3; done
done
PTB200_LIB=build/variants/q16.so python tools/express_sweep.py c4 64 -1
} > $O/r2_run26.log 2>&1
cat $O/r2_run26.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c3 64" "c4 64" "c5 16"; do
  timeout 300 python tools/variant_time.py build/variants/nostash.so $c 3
  timeout 300 python tools/variant_time.py build/variants/stash.so $c 3
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
} > $O/r2_run57.log 2>&1
cat $O/r2_run57.log

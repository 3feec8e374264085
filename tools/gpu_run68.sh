#!/bin/bash
# short rounds: a ray's chunk boxes in blocks of 4 / 2 instead of 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c5 16"; do for v in head fb4 fb2 head fb4 fb2; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 5
done; done
PT_PHASE_TIMING=1 timeout 200 python tools/phase_compare.py build/variants/fb4_phx.so c1 2>&1 | grep -v "^desc"
} > $O/r2_run68.log 2>&1
cat $O/r2_run68.log

#!/bin/bash
# CTA sizes below 768 threads (more registers per thread, fewer spills): 640, 704, and 448 threads with two rays each
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64"; do for v in base t704 t640 t448r2; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 4
done; done
} > $O/r2_run62.log 2>&1
cat $O/r2_run62.log

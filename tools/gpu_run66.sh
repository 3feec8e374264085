#!/bin/bash
# items scanned in place per workload: the build before the demand-following shares, the current build, a 6 144-item list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c3 32" "c4 16" "c5 16"; do for v in build/variants/base.so path_tracer_b200/lib/libptb200.so build/variants/items6k.so; do
  timeout 300 python tools/inplace_items.py $v $c
done; done
} > $O/r2_run66.log 2>&1
cat $O/r2_run66.log

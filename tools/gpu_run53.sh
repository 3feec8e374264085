#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in path_tracer_b200/lib/libptb200.so build/variants/b2.so build/variants/b3.so; do
  for c in "c1 100" "c2 64" "c3 64" "c4 32" "c5 16"; do timeout 300 python tools/variant_time.py $v $c 3; done
done
} > $O/r2_run53.log 2>&1
cat $O/r2_run53.log

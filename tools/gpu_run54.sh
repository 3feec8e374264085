#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in path_tracer_b200/lib/libptb200.so build/variants/ep48.so build/variants/ep96.so build/variants/fr96.so build/variants/fr256.so build/variants/fb4.so build/variants/fb16.so path_tracer_b200/lib/libptb200.so; do
  timeout 300 python tools/variant_time.py $v c1 100 5
done
} > $O/r2_run54.log 2>&1
cat $O/r2_run54.log

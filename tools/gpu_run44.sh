#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in path_tracer_b200/lib/libptb200.so build/variants/mp5.so build/variants/msc.so build/variants/mall.so; do python tools/variant_time.py $v c1 100 4; done
for v in path_tracer_b200/lib/libptb200.so build/variants/mall.so; do python tools/variant_time.py $v c2 64 3; python tools/variant_time.py $v c4 64 3; done
} > $O/r2_run44.log 2>&1
cat $O/r2_run44.log

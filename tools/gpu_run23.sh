#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in path_tracer_b200/lib/libptb200.so build/variants/u2.so build/variants/u8.so build/variants/p512.so build/variants/p768.so build/variants/s4.so build/variants/s8.so; do python tools/variant_time.py $v c4 16 3; done
} > $O/r2_run23.log 2>&1
cat $O/r2_run23.log

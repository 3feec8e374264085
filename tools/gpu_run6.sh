#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run6.log; : > $L
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/variant_time.py path_tracer_b200/lib/libptb200.so c1 8 1 > gpurun_out/r2_memcheck6.log 2>&1; tail -4 gpurun_out/r2_memcheck6.log >> $L; bash tools/gpu_run5.sh
grep -v "^=========     Host Frame\|^=========         in \|^=========                in" gpurun_out/r2_memcheck6.log | head -60 >> $L
cat $L

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o gpurun_out/r2_c4_bfs -f \
  python tools/variant_time.py path_tracer_b200/lib/libptb200.so c4 8 1 > gpurun_out/r2_ncu_c4b.log 2>&1
tail -2 gpurun_out/r2_ncu_c4b.log
timeout 300 python tools/variant_time.py path_tracer_b200/lib/libptb200.so c4 16 3

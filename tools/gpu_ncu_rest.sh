#!/bin/bash
# full ncu captures of the frame's launch of the other three workloads (reduced spp)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
LIB=path_tracer_b200/lib/libptb200.so
for c in "c2 16" "c3 16" "c5 4"; do set -- $c
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2_$1_final -f python tools/variant_time.py $LIB $1 $2 1 > $O/r2_ncu_$1_final.log 2>&1
done
ls -la $O/r2_c2_final.ncu-rep $O/r2_c3_final.ncu-rep $O/r2_c5_final.ncu-rep

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
python tools/small_region.py 352 180 96 64 100 > $O/r2_run50.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2_c1_short -f python tools/small_region.py 352 180 96 64 100 >> $O/r2_run50.log 2>&1
cat $O/r2_run50.log | tail -12

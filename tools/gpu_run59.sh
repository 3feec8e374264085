#!/bin/bash
# the two-stream test; full ncu capture of the config-5 frame launch (8 spp: with fewer there is no cost probe to skip)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
LIB=path_tracer_b200/lib/libptb200.so
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "two_streams or device_resident" 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2_c5_final -f python tools/variant_time.py $LIB c5 8 1 > $O/r2_ncu_c5_final.log 2>&1
tail -3 $O/r2_ncu_c5_final.log; ls -la $O/r2_c5_final.ncu-rep

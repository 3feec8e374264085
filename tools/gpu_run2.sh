#!/bin/bash
# round-2 GPU session 2: where did C1 lose 8 %?  + ncu of C4 and C1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in r1 head nocoop notrees r1 head; do
  timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> gpurun_out/r2_variants.log 2>&1
done
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "resume" > gpurun_out/r2_pytest2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o gpurun_out/r2_c4_wave -f \
  python tools/variant_time.py path_tracer_b200/lib/libptb200.so c4 8 1 > gpurun_out/r2_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o gpurun_out/r2_c1_wave -f \
  python tools/variant_time.py path_tracer_b200/lib/libptb200.so c1 100 1 > gpurun_out/r2_ncu_c1.log 2>&1
cat gpurun_out/r2_variants.log; tail -3 gpurun_out/r2_pytest2.log; tail -2 gpurun_out/r2_ncu_c4.log

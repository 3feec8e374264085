#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r2_run19.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest19.log 2>&1
tail -25 gpurun_out/r2_pytest19.log >> $L
for v in g1 y3 r1; do timeout 120 python tools/variant_time.py build/variants/$v.so c1 100 5 >> $L 2>&1; done
python tools/save_render.py c3 64 gpurun_out/c3_gpu_64.npy >> $L 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
cat $L

#!/bin/bash
# pixel coordinates kept in the pool (no queue_pixel() -- tile order from global memory -- at every path end): A/B, express phase cycles, parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c5 16"; do for v in head pxy head pxy; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 5
done; done
for c in "c3 64" "c4 32"; do for v in head pxy; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 4
done; done
PT_PHASE_TIMING=1 timeout 200 python tools/phase_compare.py build/variants/pxy_phx.so c1 2>&1 | grep -v "^desc"
PTB200_LIB=$PWD/build/variants/pxy.so timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
} > $O/r2_run67.log 2>&1
cat $O/r2_run67.log

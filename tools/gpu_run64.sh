#!/bin/bash
# item-list shares that follow the demand (instead of a fixed 3/8 : 5/8): A/B on every workload, parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c3 64" "c4 32" "c5 16"; do for v in base dyn; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 5
done; done
PTB200_LIB=$PWD/build/variants/dyn.so timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
} > $O/r2_run64.log 2>&1
cat $O/r2_run64.log

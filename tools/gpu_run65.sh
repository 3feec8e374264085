#!/bin/bash
# item-list shares: the fixed split while both kinds fit, by demand otherwise (dyn2): A/B, parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c5 16"; do for v in base dyn2 base dyn2; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 5
done; done
for c in "c3 64" "c4 32"; do for v in base dyn2; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 4
done; done
PTB200_LIB=$PWD/build/variants/dyn2.so timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
} > $O/r2_run65.log 2>&1
cat $O/r2_run65.log

#!/bin/bash
# the kinds' lists shaded as one list when per-kind units would need a second pass: A/B, phase cycles, parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in base pack base pack; do timeout 200 python tools/variant_time.py build/variants/$v.so c1 100 5; done
PT_PHASE_TIMING=1 timeout 200 python tools/phase_compare.py build/variants/base_ph.so c1 2>&1 | grep -v "^desc"
PT_PHASE_TIMING=1 timeout 200 python tools/phase_compare.py build/variants/pack_ph.so c1 2>&1 | grep -v "^desc"
for c in "c2 64" "c3 64" "c4 32" "c5 16"; do
  timeout 300 python tools/variant_time.py build/variants/base.so $c 3
  timeout 300 python tools/variant_time.py build/variants/pack.so $c 3
done
PTB200_LIB=$PWD/build/variants/pack.so timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
} > $O/r2_run60.log 2>&1
cat $O/r2_run60.log

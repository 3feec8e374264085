#!/bin/bash
# narrow shading units in short rounds: A/B against the committed build, phase cycles of the express rounds, pool sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in base narrow base narrow narrow_p48 narrow_p32; do timeout 200 python tools/variant_time.py build/variants/$v.so c1 100 5; done
PT_PHASE_TIMING=1 timeout 200 python tools/phase_compare.py build/variants/base_phx.so c1 2>&1 | grep -v "^desc"
PT_PHASE_TIMING=1 timeout 200 python tools/phase_compare.py build/variants/narrow_phx.so c1 2>&1 | grep -v "^desc"
for e in 12 14 17 20 24; do echo "== narrow express=$e"; PTB200_LIB=$PWD/build/variants/narrow.so timeout 200 python tools/timeline.py 100 0 1 $e | tail -2; done
for c in "c2 64" "c3 64" "c4 32" "c5 16"; do
  timeout 300 python tools/variant_time.py build/variants/base.so $c 3
  timeout 300 python tools/variant_time.py build/variants/narrow.so $c 3
done
} > $O/r2_run58.log 2>&1
cat $O/r2_run58.log

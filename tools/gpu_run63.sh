#!/bin/bash
# CTA sizes: 800 / 768 / 704 / 640 threads on every workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c3 64" "c4 32" "c5 16"; do for v in base t800 t768 t704 t640; do
  timeout 300 python tools/variant_time.py build/variants/$v.so $c 4
done; done
for v in t800 t768 t640; do for e in 17 20 24; do echo "== $v express=$e"; PTB200_LIB=$PWD/build/variants/$v.so timeout 200 python tools/timeline.py 100 0 1 $e | tail -2; done; done
} > $O/r2_run63.log 2>&1
cat $O/r2_run63.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for i in 1 2; do
python tools/variant_time.py build/variants/p5exact.so c1 100 4
python tools/variant_time.py path_tracer_b200/lib/libptb200.so c1 100 4
done
timeout 900 python -m pytest tests -m gpu -x -q -k "glibc or math or c1 or scene_parity" 2>&1 | tail -3
} > $O/r2_run49.log 2>&1
cat $O/r2_run49.log

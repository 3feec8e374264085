#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in path_tracer_b200/lib/libptb200.so build/variants/u2.so build/variants/u8.so build/variants/u16.so; do python tools/variant_time.py $v c4 64 3; done
python tools/express_sweep.py c4 64 -1 2 4 6
} > $O/r2_run52.log 2>&1
cat $O/r2_run52.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --workload c4 > $O/r2_final_c4.json 2> $O/r2_final_c4.err; head -c 200 $O/r2_final_c4.json; echo
python tools/variant_time.py path_tracer_b200/lib/libptb200.so c3 64 3
python tools/variant_time.py path_tracer_b200/lib/libptb200.so c1 100 3
} > $O/r2_run43.log 2>&1
cat $O/r2_run43.log

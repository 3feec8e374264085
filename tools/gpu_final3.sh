#!/bin/bash
# bench lines of the final build (1 GPU): c1 (default flags), c2, c3, c5 (c4: tools/gpu_run43.sh)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python bench.py > $O/r2_final_c1.json 2> $O/r2_final_c1.err
for c in c2 c3 c4 c5; do timeout 900 python bench.py --workload $c > $O/r2_final_$c.json 2> $O/r2_final_$c.err; done
for f in c1 c2 c3 c4 c5; do echo "== $f"; head -c 260 $O/r2_final_$f.json; echo; tail -n 1 $O/r2_final_$f.err; done

#!/bin/bash
# last check of the tree as committed: GPU tests, smoke(), the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>/dev/null | head -c 600; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 400; echo
} > $O/r2_verify.log 2>&1
cat $O/r2_verify.log

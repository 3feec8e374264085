#!/bin/bash
# quick sanitizer passes of the last kernel change (item-list shares): default scene + RTIOW, tiny
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "--- memcheck"; timeout 45 compute-sanitizer --tool memcheck python tools/tiny_render2.py 2>&1 | grep -v "^=========\s*$" | tail -4
echo "--- synccheck"; timeout 45 compute-sanitizer --tool synccheck python tools/tiny_render2.py 2>&1 | grep -v "^=========\s*$" | tail -4
} > gpurun_out/r2b_memsync.log 2>&1
cat gpurun_out/r2b_memsync.log

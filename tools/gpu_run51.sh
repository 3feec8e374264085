#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- racecheck (wavefront kernel)"; PT_TINY_LPT=1 timeout 1500 compute-sanitizer --tool racecheck python tools/tiny_render.py 2>&1 | grep -v "^=========\s*$" | tail -16
echo "--- memcheck (wavefront kernel)"; PT_TINY_LPT=1 timeout 1500 compute-sanitizer --tool memcheck python tools/tiny_render.py 2>&1 | tail -5
echo "--- synccheck (wavefront kernel)"; timeout 1200 compute-sanitizer --tool synccheck python tools/tiny_render.py 2>&1 | tail -2
echo "--- memcheck (lane kernel)"; timeout 1200 compute-sanitizer --tool memcheck python tools/tiny_render.py 1 2>&1 | tail -2
} > $O/r2_final_sanitizer.log 2>&1
cat $O/r2_final_sanitizer.log

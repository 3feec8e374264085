#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12 > $O/r2_run56.log
cat $O/r2_run56.log

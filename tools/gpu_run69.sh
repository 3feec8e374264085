#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "shares_follow" 2>&1 | grep -E "^E " | head -5
PTB200_LIB=$PWD/build/variants/base.so timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "shares_follow" 2>&1 | grep -E "^E " | head -5
for v in path_tracer_b200/lib/libptb200.so build/variants/base.so; do timeout 200 python tools/inplace_items.py $v c2 8; done

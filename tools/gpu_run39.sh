#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PTB200_LIB=build/variants/planes.so python tools/express_sweep.py c4 64 -1 4 17
python tools/variant_time.py build/variants/planes.so c4 256 2
python tools/variant_time.py build/variants/planes.so c1 100 3
python tools/variant_time.py build/variants/planes.so c3 64 3
python tools/variant_time.py build/variants/planes.so c2 64 3
python tools/variant_time.py build/variants/planes.so c5 16 3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > $O/r2_run39.log 2>&1
cat $O/r2_run39.log

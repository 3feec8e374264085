#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
PTB200_LIB=build/variants/spill.so python tools/tail_probe.py c4 64
python tools/variant_time.py build/variants/spill.so c4 256 2
python tools/variant_time.py build/variants/spill.so c1 100 3
python tools/variant_time.py build/variants/spill.so c3 64 3
PTB200_LIB=build/variants/spill.so timeout 900 python -m pytest tests -m gpu -x -q -k "flat or config4 or mesh or tree or overflow or table" 2>&1 | tail -3
} > $O/r2_run35.log 2>&1
cat $O/r2_run35.log

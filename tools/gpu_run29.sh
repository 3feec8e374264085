#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
python tools/variant_time.py build/variants/q16.so c4 64 3
python tools/variant_time.py build/variants/adapt.so c4 64 3
python tools/variant_time.py build/variants/adapt.so c4 16 3
python tools/variant_time.py build/variants/adapt.so c1 100 3
python tools/variant_time.py build/variants/adapt.so c3 64 3
PT_PHASE_TIMING=1 python tools/phase_compare.py build/variants/adaptph.so c4 32
PTB200_LIB=build/variants/adapt.so python tools/express_sweep.py c4 64 -1 4 17 32
PTB200_LIB=build/variants/adapt.so timeout 900 python -m pytest tests -m gpu -x -q -k "flat or config4 or mesh or tree" 2>&1 | tail -3
} > $O/r2_run29.log 2>&1
cat $O/r2_run29.log

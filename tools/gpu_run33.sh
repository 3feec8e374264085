#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for v in pr4 pr8; do PTB200_LIB=build/variants/$v.so python tools/express_sweep.py c4 64 -1 17; done
PTB200_LIB=build/variants/prauto.so python tools/express_sweep.py c4 256 -1
PTB200_LIB=build/variants/prauto.so python tools/express_sweep.py c3 1024 -1
PTB200_LIB=build/variants/adapt.so python tools/express_sweep.py c3 1024 -1
} > $O/r2_run33.log 2>&1
cat $O/r2_run33.log

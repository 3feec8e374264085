#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/save_render.py c3 64 gpurun_out/c3_gpu_64.npy
PTB200_LIB=build/variants/r1.so python tools/save_render.py c3 64 gpurun_out/c3_r1_64.npy 2>&1 | tail -1

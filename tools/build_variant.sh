#!/bin/bash
# build_variant.sh NAME [-Dflags...] : libptb200 with extra defines -> build/variants/NAME.so (experiments)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-w -Iinclude -Ipath_tracer_b200/csrc "$@" -shared -cudart static -o build/variants/$name.so \
  path_tracer_b200/csrc/pt_wave.cu path_tracer_b200/csrc/pt_lane.cu path_tracer_b200/csrc/pt_api.cu path_tracer_b200/csrc/pt_pack.cpp
echo built build/variants/$name.so

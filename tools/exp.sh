#!/bin/bash
# usage: tools/exp.sh lib1.so lib2.so ...   -> times C1 at 100 spp with each build
for l in "$@"; do echo "== $l"; PTB200_LIB=$PWD/path_tracer_b200/lib/$l python tools/run_c1.py 100 3 | tail -2; done

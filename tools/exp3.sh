#!/bin/bash
# usage: tools/exp3.sh "express counts" lib1.so lib2.so ...  -> timeline of C1 at 100 spp per build and express CTA count
ex="$1"; shift
for l in "$@"; do for e in $ex; do echo "== $l express=$e"; PTB200_LIB=$PWD/path_tracer_b200/lib/$l python tools/timeline.py 100 0 1 $e | tail -2; done; done

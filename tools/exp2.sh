#!/bin/bash
for l in "$@"; do echo "== $l"; PTB200_LIB=$PWD/path_tracer_b200/lib/$l python tools/timeline.py 100 | tail -2; done

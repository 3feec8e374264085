#!/bin/bash
# shading units shared between kinds: PT_PACK_MODE 0 (none) / 1 (all, rounds without emitters and media) / 2 (bg..dielectric) / 3 (all, cheap kinds first)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
for c in "c1 100" "c2 64" "c3 64" "c4 32" "c5 16"; do for m in 0 1 2 3; do
  timeout 300 python tools/variant_time.py build/variants/pm$m.so $c 4
done; done
} > $O/r2_run61.log 2>&1
cat $O/r2_run61.log

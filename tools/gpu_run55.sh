#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, time
sys.path.insert(0, "tests")
import scenes
from path_tracer_b200 import render as R
sc, cam, _ = scenes.load_c1()
t0 = time.perf_counter(); R.render_single_task(sc, cam, 64, 48, 4, 50); dt = time.perf_counter() - t0
st = R.stats()
print("single task c1 64x48x4: %.2f s, kernel %.1f ms, %d scans -> %.2f us per scan" % (dt, st["kernel_ms"], st["scans"], 1e3 * st["kernel_ms"] / st["scans"]))
PY
} > $O/r2_run55.log 2>&1
cat $O/r2_run55.log

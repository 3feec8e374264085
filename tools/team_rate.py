import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from path_tracer_b200 import render as R
sc, cam, (w, h, _, d) = scenes.load_c1()
L = R.lib(); L.pt_debug_set_kernel(1)
spp = 10
for t in (1, 2, 4, 8, 16, 32):
    L.pt_debug_set_team_size(t)
    for rep in range(2):
        R.render(sc, cam, w, h, spp, d); st = R.stats()
    warps = 148 * 5 * 4
    rounds_per_warp = st["scans"] * t / 32 / warps
    print("team %2d: kernel %.2f ms, scans %d -> %.1f Mpaths/s, %.2f us per warp-round (%.0f rounds per warp)" % (
        t, st["kernel_ms"], st["scans"], w * h * spp / st["kernel_ms"] / 1e3, st["kernel_ms"] * 1e3 / rounds_per_warp, rounds_per_warp))

"""Kernel time of the BASELINE configurations' scenes (tests/scenes.py builders) with and without chunk culling."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from path_tracer_b200 import render as R
L = R.lib()
cases = [("C1 default scene 800x480x100", scenes.load_c1()[:2], 800, 480, 100),
         ("C2 RTIOW 1920x1080x64", scenes.rtiow(16 / 9), 1920, 1080, 64),
         ("C3 Cornell 1024x1024x64 (of 1024 spp)", scenes.cornell(1.0), 1024, 1024, 64),
         ("C4-like 248 triangles 1920x1080x16", scenes.triangle_mesh(16 / 9), 1920, 1080, 16),
         ("C5 motion blur 3840x2160x8 (of 4096 spp)", scenes.motion_blur(16 / 9), 3840, 2160, 8)]
for name, (sc, cam), w, h, spp in cases:
    row = []
    for cull in (1, 0):
        L.pt_debug_set_cull(cull)
        R.render(sc, cam, w, h, spp, 50)
        R.render(sc, cam, w, h, spp, 50)
        st = R.stats()
        row.append(w * h * spp / st["kernel_ms"] / 1e3)
    L.pt_debug_set_cull(1)
    print("%-44s %8.1f Mpaths/s (without chunk culling %8.1f)" % (name, row[0], row[1]), flush=True)

"""Render C1 (default scene) once or a few times through the device-resident API; used under ncu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from path_tracer_b200 import render as R
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc, cam, meta = scenes.load_c1()
for i in range(reps):
    t = time.time(); a = R.render(sc, cam, 800, 480, spp, 50); dt = time.time() - t
    st = R.stats()
    print("C1 spp %d: wall %.3f s kernel %.2f ms -> %.1f Mpaths/s kernel" % (spp, dt, st["kernel_ms"], 800*480*spp/st["kernel_ms"]/1e3), flush=True)

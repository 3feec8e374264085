"""A few tiny renders through the C-ABI (for compute-sanitizer runs).  usage: tiny_render.py [kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from path_tracer_b200 import render as R
if len(sys.argv) > 1: R.lib().pt_debug_set_kernel(int(sys.argv[1]))
sc, cam, _ = scenes.load_c1()
a = R.render(sc, cam, 96, 64, 3, 50)
print("c1", float(a.mean()), R.stats()["scans"])
if os.environ.get("PT_TINY_LPT"):  # large enough for the cost probe, the LPT order, express CTAs and hand-offs
    a = R.render(sc, cam, 320, 192, 12, 50)
    print("c1 with LPT order and express CTAs", float(a.mean()), R.stats()["scans"], R.stats()["kernel_launches"])
for name in ("media", "moving", "shapes"):
    s, c = scenes.ALL[name](4 / 3)
    a = R.render(s, c, 48, 36, 2, 50)
    print(name, float(a.mean()), R.stats()["scans"])
